O=gpurun_out
for w in c3_coord_2d c4_coord_3d; do
 for g in 32 64 128; do
  EDM_L2_FETCH=$g python bench.py --workload $w --steps 10 --no-cpu-baseline > $O/r2s_l2f${g}_${w}.json 2>> $O/r2s_err.log
 done
done
EDM_L2_FETCH=32 python bench.py --steps 10 --no-cpu-baseline > $O/r2s_l2f32_c2.json 2>> $O/r2s_err.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2s_*.json")):
    l=[x for x in open(f) if x.startswith("{")]
    if not l: print(f,"EMPTY"); continue
    d=json.loads(l[-1]); print(f, "%.4f ms"%d["ms_per_step"], "kernel %.4f"%d["roofline"]["kernel_ms"])
PY
tail -3 $O/r2s_err.log
