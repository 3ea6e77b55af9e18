timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cmvn" 2>&1 | tail -1
python bench.py --config p_cmvn_long --steps 20 > gpurun_out/r2z_bench_p_cmvn_long.json 2> gpurun_out/r2z_bench_p_cmvn_long.err
python -c "import json; d=json.loads(open('gpurun_out/r2z_bench_p_cmvn_long.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'])"
ncu --set full --import-source on --clock-control none -k regex:k_cmvn_coop -c 1 -f -o gpurun_out/prof_cmvn_coop_r2 python bench.py --config p_cmvn_long --steps 1 --warmup 1 > gpurun_out/ncu_r2_cmvn_long.log 2>&1
