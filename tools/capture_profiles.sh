#!/bin/bash
# Evidence run on the GPU box (gpurun -- bash tools/capture_profiles.sh TAG): bench lines first (never under a
# profiler), then ncu launch lists, then one `ncu --set full` capture per kernel of interest.  Everything lands in
# gpurun_out/; tools/condense_profiles.sh TAG condenses them into profiles/ back in the container.
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
python bench.py --csrc-hash > $O/${TAG}_csrc_hash.txt
NCU="ncu --set full --clock-control none --import-source on"
LIST="ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv"
# 1. bench lines
python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_c2_n1.json 2>> $O/${TAG}_err.log
python bench.py --impl reference --steps 10 --warmup 2 > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_err.log
for w in c1_coord_1d c3_coord_2d c4_coord_3d c5_pair_rdf_backlog c2_pair_rdf_local_tempering; do
  python bench.py --workload $w --steps 20 --no-cpu-baseline > $O/${TAG}_bench_${w}_n1.json 2>> $O/${TAG}_err.log
done
for w in c3_coord_2d c4_coord_3d; do
  for ord in cell strip; do
    python bench.py --workload $w --input-order $ord --steps 10 --no-cpu-baseline > $O/${TAG}_bench_${w}_${ord}_n1.json 2>> $O/${TAG}_err.log
  done
done
# 2. launch lists
$LIST --log-file $O/${TAG}_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
for w in c1_coord_1d c3_coord_2d c4_coord_3d; do
  $LIST --log-file $O/${TAG}_launches_${w}.csv python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
done
# 3. full captures, one launch each
$NCU -k regex:block_eval_kernel -s 6 -c 1 -o $O/${TAG}_block_eval -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
$NCU -k regex:block_find_kernel -s 6 -c 1 -o $O/${TAG}_block_find -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
$NCU -k regex:forces_kernel -s 6 -c 1 -o $O/${TAG}_forces_c3 -f python bench.py --workload c3_coord_2d --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
$NCU -k regex:forces_kernel -s 6 -c 1 -o $O/${TAG}_forces_c3_cell -f python bench.py --workload c3_coord_2d --input-order cell --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
$NCU -k regex:forces_kernel -s 6 -c 1 -o $O/${TAG}_forces_c4 -f python bench.py --workload c4_coord_3d --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
$NCU -k regex:forces_kernel -s 6 -c 1 -o $O/${TAG}_forces_c4_cell -f python bench.py --workload c4_coord_3d --input-order cell --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
$NCU -k regex:round_plan_kernel -s 6 -c 1 -o $O/${TAG}_plan_c3 -f python bench.py --workload c3_coord_2d --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
$NCU -k regex:round_integrals_kernel -s 6 -c 1 -o $O/${TAG}_integrals_c4 -f python bench.py --workload c4_coord_3d --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
$NCU -k regex:round_deposit_kernel -s 6 -c 1 -o $O/${TAG}_deposit_c4 -f python bench.py --workload c4_coord_3d --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
$NCU -k regex:deposit1d_owner_kernel -c 1 -o $O/${TAG}_deposit1d -f python tools/prof_deposit.py > /dev/null 2>&1
ls -la $O | grep ${TAG} | awk '{print $5, $9}'
tail -3 $O/${TAG}_err.log
