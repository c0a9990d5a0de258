# A/B of the hill-exchange transports on N GPUs (gpurun --gpus N -- bash tools/experiments/peer_ab.sh TAG N)
O=gpurun_out
T=${1:-r2v}
N=${2:-2}
timeout 900 python -m pytest tests/test_multi_rank.py tests/test_gpu_exchange.py -m gpu -x -q > $O/${T}_tests.log 2>&1
tail -5 $O/${T}_tests.log
P=29500
for w in c3_coord_2d c4_coord_3d c2_pair_rdf; do
  for nop2p in 0 1; do
    P=$((P+1))
    EDM_B200_NO_P2P=$nop2p timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P \
      bench.py --gpus $N --workload $w --steps 20 --no-cpu-baseline > $O/${T}_${w}_n${N}_nop2p${nop2p}.json 2>> $O/${T}_err.log
  done
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${T}_*.json")):
    l=[x for x in open(f) if x.startswith("{")]
    if not l: print(f,"EMPTY"); continue
    d=json.loads(l[-1]); print(f.split("/")[-1], "%.4f ms"%d["ms_per_step"], d.get("step_breakdown_ms",{}).get("hill_round"), (d.get("timing") or {}).get("exchange","")[:60])
PY
grep -v "^\*\*\*\|^$\|OMP_NUM" $O/${T}_err.log | tail -5
