python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline'], d['e2e'], d['clocks'])"
ncu --set full --clock-control none --import-source on -k regex:k_aop_scalar_tiled -s 3 -c 1 -o gpurun_out/prof_aop_scalar_tiled python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/b2.log 2>&1
