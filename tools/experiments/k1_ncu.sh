O=gpurun_out
T=${1:-r2w}
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:forces_kernel -s 6 -c 1 -o $O/${T}_forces_c3_cell -f python bench.py --workload c3_coord_2d --input-order cell --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
$NCU -k regex:forces_kernel -s 6 -c 1 -o $O/${T}_forces_c4_cell -f python bench.py --workload c4_coord_3d --input-order cell --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la $O | grep ${T}
