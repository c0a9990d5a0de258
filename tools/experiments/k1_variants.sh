# K1 variant sweep (EDM_K1 experiment hook in edm_bias_update_forces_dev): bench lines per variant and input order
O=gpurun_out
T=${1:-r2t}
run() { # workload order variant
  EDM_K1=$3 python bench.py --workload $1 --input-order $2 --steps 10 --no-cpu-baseline > $O/${T}_$1_$2_$3.json 2>> $O/${T}_err.log
}
for ord in random cell; do
  for v in base u2 u2b3 u1 u1b4 pipe pipeb2 pipeb4; do run c3_coord_2d $ord $v; done
  for v in base u2 u1 u1b3 pipe pipeb3; do run c4_coord_3d $ord $v; done
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${T}_*.json")):
    l=[x for x in open(f) if x.startswith("{")]
    if not l: print(f,"EMPTY"); continue
    d=json.loads(l[-1]); print(f.split("/")[-1], "%.4f ms"%d["ms_per_step"], "kernel %.4f"%d["roofline"]["kernel_ms"])
PY
tail -3 $O/${T}_err.log
