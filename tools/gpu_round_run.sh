#!/bin/bash
# One-GPU measurement pass of a round: tests, bench lines of every BASELINE config, ncu launch list / traffic / full captures.
# Run on the GPU box through gpurun; everything lands in gpurun_out/ with the prefix given as $1.
P=${1:-r02}
O=gpurun_out
mkdir -p $O
(ls -la baseline/_ref 2>&1 | head -3; nvidia-smi -L; nproc; nvidia-smi --query-gpu=clocks.max.sm,power.limit --format=csv) > $O/${P}_gpu_env.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -s > $O/${P}_pytest.log 2>&1; echo "pytest rc $?" >> $O/${P}_pytest.log
timeout 400 python bench.py > $O/${P}_bench_infer_n1.json 2> $O/${P}_bench.err
timeout 200 python bench.py --impl reference > $O/${P}_bench_ref.json 2>> $O/${P}_bench.err
B="--no-cpu-baseline --no-torch-gpu-baseline"
timeout 200 python bench.py --workload predictive $B > $O/${P}_bench_pred.json 2>> $O/${P}_bench.err
for n in 1 10 30; do timeout 300 python bench.py --bridge-steps $n $B > $O/${P}_bench_N$n.json 2>> $O/${P}_bench.err; done
timeout 300 python bench.py --seconds 30 --utts 32 --micro-batch 16 $B > $O/${P}_bench_30s.json 2>> $O/${P}_bench.err
timeout 300 python bench.py --workload files > $O/${P}_bench_files.json 2>> $O/${P}_bench.err
timeout 300 python bench.py --workload train > $O/${P}_bench_train_n1.json 2>> $O/${P}_bench.err
timeout 400 python bench.py --workload tfgridnet --utts 72 > $O/${P}_bench_tfgridnet.json 2>> $O/${P}_bench.err
timeout 300 python bench.py --workload tfgridnet_predictive --utts 72 > $O/${P}_bench_tfgridnet_pred.json 2>> $O/${P}_bench.err
# ncu: launch list of one timed step (128 utterances = one micro-batch), conv DRAM traffic of one forward, full captures
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1270 --launch-count 1270 --csv --log-file $O/${P}_launches.csv \
  python bench.py --utts 128 --steps 1 --warmup 3 $B > $O/${P}_ncu_launches.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:conv_igemm --launch-skip 228 --launch-count 114 --csv --log-file $O/${P}_conv_traffic.csv \
  python tools/forward_profile.py 128 256 > $O/${P}_ncu_traffic.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_igemm --launch-skip 232 --launch-count 6 -o $O/${P}_conv_full \
  python tools/forward_profile.py 128 256 > $O/${P}_ncu_conv_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lstm_sweep_tc --launch-skip 4 --launch-count 1 -o $O/${P}_lstm_full \
  python tools/tfg_bench.py 16 > $O/${P}_ncu_lstm_full.log 2>&1
timeout 300 python tools/backward_profile.py 16 > $O/${P}_bwd_profile.txt 2>&1
timeout 200 python tools/tfg_bench.py 4 16 32 > $O/${P}_tfg_forward.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${P}_tfg_launches.csv python tools/tfg_bench.py 16 > $O/${P}_ncu_tfg.log 2>&1
tail -3 $O/${P}_pytest.log
