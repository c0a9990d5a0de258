# N-GPU lines of every multi-GPU workload: gpurun --gpus N -- bash tools/experiments/scale_n.sh TAG N
O=gpurun_out
T=${1:-r02c}
N=${2:-8}
P=29600
run() { # workload extra-env-name extra-env-value suffix
  P=$((P+1))
  env $2=$3 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P \
    bench.py --gpus $N --workload $1 --steps 20 --warmup 5 --no-cpu-baseline 2>> $O/${T}_err.log | grep '^{' > $O/${T}_bench_$1_n${N}$4.json
}
run c2_pair_rdf EDM_X 0 ""
run c3_coord_2d EDM_X 0 ""
run c4_coord_3d EDM_X 0 ""
run c2_one_box EDM_X 0 ""
run c3_coord_2d EDM_B200_NO_P2P 1 _nccl
run c2_pair_rdf EDM_B200_NO_P2P 1 _nccl
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${T}_bench_*.json")):
    l=[x for x in open(f) if x.startswith("{")]
    if not l: print(f,"EMPTY"); continue
    d=json.loads(l[-1]); print(f.split("/")[-1], "%.4f ms"%d["ms_per_step"], "%.4g"%d["value"], "e2e %.4g"%d.get("e2e",{}).get("value",0), d.get("step_breakdown_ms",{}).get("hill_round"))
PY
grep -v "^\*\*\*\|^$\|OMP_NUM" $O/${T}_err.log | tail -5
