python bench.py --steps 3 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_rt_v1.json; cat gpurun_out/bench_rt_v1.json | cut -c1-2500
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r01_rt_v1.csv python bench.py --steps 1 --warmup 3 --nobs 512 --no-cpu --no-aop > gpurun_out/b.log 2>&1
tail -3 gpurun_out/b.log | cut -c1-300
