#!/bin/bash
# Back in the container after `gpurun -- bash tools/capture_profiles.sh TAG`: condenses gpurun_out/TAG_* into profiles/
# (bench lines and launch lists copied, every .ncu-rep summarised, profiles/ncu_roofline.json refreshed under the
# csrc hash the capture ran on).
TAG=${1:-r02}
O=gpurun_out
P=profiles
H=$O/${TAG}_csrc_hash.txt   # ncu_summary.py picks the line of the kernel's translation unit
cp $O/${TAG}_bench_*.json $O/${TAG}_launches_*.csv $P/ 2>/dev/null
sum() { # capture-name [workload kernel]...
  rep=$O/${TAG}_$1.ncu-rep
  [ -f $rep ] || { echo "missing $rep"; return; }
  out=$P/${TAG}_$1_ncu_selected.txt
  shift
  if [ $# -ge 2 ]; then
    python tools/ncu_summary.py $rep --roofline "$1" "$2" $H $out > $out
    shift 2
    while [ $# -ge 2 ]; do python tools/ncu_summary.py $rep --roofline "$1" "$2" $H $out > /dev/null; shift 2; done
  else
    python tools/ncu_summary.py $rep > $out
  fi
  python tools/ncu_hot_lines.py $rep 16 > ${out%_ncu_selected.txt}_hot_lines.txt 2>/dev/null
}
sum block_eval c2_pair_rdf block_eval_kernel c5_pair_rdf_backlog block_eval_kernel c2_pair_rdf_local_tempering block_eval_kernel c2_one_box block_eval_kernel
sum block_find
sum forces_c3 c3_coord_2d "forces_kernel<2>"
sum forces_c3_cell "c3_coord_2d@cell" "forces_kernel<2>"
sum forces_c4 c4_coord_3d "forces_kernel<3>"
sum forces_c4_cell "c4_coord_3d@cell" "forces_kernel<3>"
sum plan_c3
sum integrals_c4
sum deposit_c4
sum deposit1d
ls $P | grep ${TAG} | wc -l
