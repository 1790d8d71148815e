python scripts/bench_prove.py 1024 1 2>&1 | tail -6
python scripts/bench_prove.py 1024 1 > gpurun_out/pp.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_prove.csv python scripts/bench_prove.py 1024 1 > gpurun_out/ncu_p.log 2>&1; tail -1 gpurun_out/ncu_p.log
