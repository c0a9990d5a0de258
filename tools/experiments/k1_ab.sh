# A/B of two builds of the library on the coordinate workloads: bash tools/experiments/k1_ab.sh TAG ALT_LIB
O=gpurun_out
T=${1:-r2x}
ALT=$2
for w in c3_coord_2d c4_coord_3d; do for ord in random cell strip; do
  python bench.py --workload $w --input-order $ord --steps 10 --no-cpu-baseline > $O/${T}_${w}_${ord}_main.json 2>> $O/${T}_err.log
  EDM_B200_LIB=$ALT python bench.py --workload $w --input-order $ord --steps 10 --no-cpu-baseline > $O/${T}_${w}_${ord}_alt.json 2>> $O/${T}_err.log
done; done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${T}_*.json")):
    l=[x for x in open(f) if x.startswith("{")]
    if not l: print(f,"EMPTY"); continue
    d=json.loads(l[-1]); print(f.split("/")[-1], "%.4f ms"%d["ms_per_step"], "kernel %.4f"%d["roofline"]["kernel_ms"])
PY
tail -3 $O/${T}_err.log
