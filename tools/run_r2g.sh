#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2g_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r2g_gputests.log
timeout 300 python tools/parity_margins.py > gpurun_out/r2g_parity_margins.json 2> gpurun_out/r2g_parity_margins.err; echo "margins rc=$?"; tail -3 gpurun_out/r2g_parity_margins.err; head -c 1500 gpurun_out/r2g_parity_margins.json
timeout 200 python tools/pcie_ceiling.py --register 2>/dev/null | tail -1 | tee gpurun_out/r2g_pcie_1gpu.json
for c in p_cmvn p_cmvn_long; do timeout 120 python bench.py --config $c --steps 10 2>/dev/null | tee gpurun_out/r2g_$c.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$c', round(d['ms_per_step'],3),'ms frac', round(d['roofline']['frac'],3))"; done
for c in c3 c4 c5 c1; do timeout 600 python bench.py --config $c --steps 10 > gpurun_out/r2g_bench_$c.json 2> gpurun_out/r2g_bench_$c.err; python -c "
import json
d=json.loads(open('gpurun_out/r2g_bench_$c.json').read().strip().splitlines()[-1])
print('$c', d['kernel'], round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), 'e2e', d['e2e'] and round(d['e2e']['value']), d['e2e'] and d['e2e'].get('frac_of_pcie_ceiling'), 'pageable', d['e2e_pageable'] and round(d['e2e_pageable']['value']), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value']))"; done
