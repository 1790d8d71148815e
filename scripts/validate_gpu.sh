#!/bin/bash
# GPU validation as run at the end of round 1 (under gpurun): GPU suite, smoke, bench (both arms), then the ncu launch lists of the verification step and a proving call
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/t.log 2>&1; tail -3 gpurun_out/t.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 500 python bench.py > gpurun_out/bench_r01_n1.json 2> gpurun_out/bench_err.log; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_err.log
timeout 300 python bench.py --impl reference > gpurun_out/bench_r01_n1_ref.json 2>> gpurun_out/bench_err.log; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01f.csv python bench.py --steps 3 --warmup 3 --lanes 1 --extras 0 > gpurun_out/ncu_f.log 2>&1; echo "ncu verify rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_prove4.csv python scripts/prove_lanes_probe.py 1024 1 > gpurun_out/ncu_p.log 2>&1; echo "ncu prove rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_r01_n1.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['one_batch_at_a_time']['value'], d['roofline']['frac'], d['roofline']['whole_step']['frac'])
print(d['extras']['prove']['value'], d['extras']['prove']['frac_of_int32_mul_peak'], {k:(round(v['mpoints_per_s'],1),v['window_bits']) for k,v in d['extras']['msm'].items()})
print(d['cpu_baseline'], d['clocks'])
P
