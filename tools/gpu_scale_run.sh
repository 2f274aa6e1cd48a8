#!/bin/bash
# Multi-GPU measurement pass (run through `gpurun --gpus N`): strong-scaling inference line, weak-scaling line, DDP training line.
# usage: tools/gpu_scale_run.sh <N> [prefix]
N=${1:-2}; P=${2:-r02}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
B="--no-cpu-baseline --no-torch-gpu-baseline"
timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 3 $B > $O/${P}_bench_infer_n$N.json 2> $O/${P}_scale_n$N.err
timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 3 --scaling weak $B > $O/${P}_bench_infer_weak_n$N.json 2>> $O/${P}_scale_n$N.err
timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 3 --workload train > $O/${P}_bench_train_n$N.json 2>> $O/${P}_scale_n$N.err
if [ "$N" = "2" ]; then timeout 300 python -m pytest tests -m gpu -q -k "device" > $O/${P}_pytest_2gpu.log 2>&1; fi
tail -c 600 $O/${P}_bench_infer_n$N.json; echo; tail -5 $O/${P}_scale_n$N.err
