#!/bin/bash
# test + bench half of a measurement pass (run through gpurun)
P=${1:-r02g}; O=gpurun_out; mkdir -p $O
(nvidia-smi -L; nproc; nvidia-smi --query-gpu=clocks.max.sm,power.limit --format=csv) > $O/${P}_gpu_env.txt 2>&1
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
timeout 120 python tools/forward_profile.py 128 256 > $O/${P}_forward_profile.txt 2>&1
tail -3 $O/${P}_pytest.log; for f in infer_n1 pred N1 N10 N30 30s files train_n1 tfgridnet tfgridnet_pred; do python -c "
import json,sys
try:
    d=json.load(open('$O/${P}_bench_$f.json')); print('$f', round(d['value'],1), d['unit'], 'e2e', round(d['e2e']['value'],1), d.get('roofline',{}).get('frac'), d.get('roofline',{}).get('traffic'), d['clocks']['sm_mhz'])
except Exception as e: print('$f FAILED', e)
"; done
