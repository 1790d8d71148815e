#!/bin/bash
# one GPU call: full GPU suite, smoke, default bench + reference arm, host-CPU probe with 4 cores
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/t.log 2>&1; tail -3 gpurun_out/t.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench_r01_n1.json 2> gpurun_out/bench_err.log; tail -c 600 gpurun_out/bench_err.log
timeout 300 python bench.py --impl reference > gpurun_out/bench_r01_n1_ref.json 2>> gpurun_out/bench_err.log
echo "== cpus=0-3" > gpurun_out/wait_probe2.log
PROBE_ONE_MODE=1 timeout 200 taskset -c 0-3 python scripts/pipeline_probe.py 48 32 >> gpurun_out/wait_probe2.log 2>&1
timeout 100 python scripts/e2e_cpu_cost.py >> gpurun_out/wait_probe2.log 2>&1
cat gpurun_out/wait_probe2.log
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_r01_n1.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['e2e']['one_call_at_a_time'], d['one_batch_at_a_time']['value'])
print(d['extras']['prove']['value'], {k:(v['mpoints_per_s'],v['window_bits']) for k,v in d['extras']['msm'].items()})
print(d['cpu_baseline'])
P
cat gpurun_out/bench_r01_n1_ref.json | head -c 600
