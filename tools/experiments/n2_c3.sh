O=gpurun_out
T=${1:-r2z}
N=${2:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --workload c3_coord_2d --steps 20 --no-cpu-baseline 2>> $O/${T}_err.log | grep '^{' > $O/${T}_c3_n$N.json
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_c3_n$N.json").read())
print(d["ms_per_step"], d["step_breakdown_ms"])
print(d["round_stamps_us"]["overlapped_step"])
print(d["round_stamps_us"]["overlapped_step_exchange"]["us"])
PY
tail -2 $O/${T}_err.log
