#!/bin/bash
# round 2, call A: the new host path (tests + bench c2 with every e2e leg)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r2a_gpu.txt; nproc >> gpurun_out/r2a_gpu.txt; free -g | head -2 >> gpurun_out/r2a_gpu.txt
timeout 900 python -m pytest tests/test_gpu_hostpath.py -x -q -m gpu > gpurun_out/r2a_hostpath.log 2>&1; echo "hostpath rc=$?"; tail -5 gpurun_out/r2a_hostpath.log
timeout 600 python bench.py --config c2 --steps 20 --warmup 5 > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err; echo "bench rc=$?"; tail -3 gpurun_out/r2a_bench_c2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2a_bench_c2.json'))
for k in ('ms_per_step','e2e','e2e_pageable','pcie_ceiling','cpu_baseline'):
    print(k, d.get(k))
PY
for t in 0 4 8 16; do SS_COPY_THREADS=$t timeout 300 python bench.py --config c2 --steps 5 --no-cpu --e2e-modes pageable 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('copy_threads $t', d['e2e_pageable']['ms_per_step'], d['e2e_pageable']['gb_per_s_per_gpu'])"; done
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2a_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -5 gpurun_out/r2a_gputests.log
