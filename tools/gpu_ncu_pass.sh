#!/bin/bash
# ncu half of a measurement pass (run through gpurun): conv DRAM traffic of one forward, launch list of one timed step, full capture
P=${1:-r02g}; O=gpurun_out; mkdir -p $O
B="--no-cpu-baseline --no-torch-gpu-baseline"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:conv_igemm --launch-skip 228 --launch-count 114 --csv --log-file $O/${P}_conv_traffic.csv \
  python tools/forward_profile.py 128 256 > $O/${P}_ncu_traffic.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1270 --launch-count 1270 --csv --log-file $O/${P}_launches.csv \
  python bench.py --utts 128 --steps 1 --warmup 3 $B > $O/${P}_ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_igemm --launch-skip 232 --launch-count 6 -o $O/${P}_conv_full \
  python tools/forward_profile.py 128 256 > $O/${P}_ncu_conv_full.log 2>&1
ls -la $O/${P}_*
