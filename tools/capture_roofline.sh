#!/bin/bash
# Short form of capture_profiles.sh after a change that only touches the roofline kernels: the default bench line, the
# C2 launch list and one `ncu --set full` capture of each kernel bench.py reports a roofline for.
TAG=${1:-r02}
WHAT=${2:-all}   # "pair": only the pair kernels (a change confined to edm_pair.cu)
O=gpurun_out
mkdir -p $O
python bench.py --csrc-hash > $O/${TAG}_csrc_hash.txt
NCU="ncu --set full --clock-control none --import-source on"
LIST="ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv"
python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_c2_n1.json 2>> $O/${TAG}_err.log
$LIST --log-file $O/${TAG}_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
$NCU -k regex:block_eval_kernel -s 6 -c 1 -o $O/${TAG}_block_eval -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
$NCU -k regex:block_find_kernel -s 6 -c 1 -o $O/${TAG}_block_find -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
if [ "$WHAT" = all ]; then
$NCU -k regex:forces_kernel -s 6 -c 1 -o $O/${TAG}_forces_c3 -f python bench.py --workload c3_coord_2d --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
$NCU -k regex:forces_kernel -s 6 -c 1 -o $O/${TAG}_forces_c3_cell -f python bench.py --workload c3_coord_2d --input-order cell --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
$NCU -k regex:forces_kernel -s 6 -c 1 -o $O/${TAG}_forces_c4 -f python bench.py --workload c4_coord_3d --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
$NCU -k regex:forces_kernel -s 6 -c 1 -o $O/${TAG}_forces_c4_cell -f python bench.py --workload c4_coord_3d --input-order cell --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
fi
ls -la $O | grep ${TAG} | awk '{print $5, $9}'
