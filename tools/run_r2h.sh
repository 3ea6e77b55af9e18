#!/bin/bash
mkdir -p gpurun_out
SS_SWEEP_BLOCKS=12 timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2h_gputests.log 2>&1; echo "gpu tests (sweep x12) rc=$?"; tail -6 gpurun_out/r2h_gputests.log
timeout 300 python tools/parity_margins.py > gpurun_out/r2h_parity_margins.json 2> gpurun_out/r2h_parity_margins.err; echo "margins rc=$?"
timeout 900 python bench.py --config c5 --steps 5 > gpurun_out/r2h_bench_c5.json 2> gpurun_out/r2h_bench_c5.err; python -c "
import json
d=json.loads(open('gpurun_out/r2h_bench_c5.json').read().strip().splitlines()[-1])
print('c5', d['kernel'], round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), 'e2e', d['e2e'], 'pageable', d['e2e_pageable'], 'cpu', d['cpu_baseline'])"
