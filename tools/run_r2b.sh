#!/bin/bash
# round 2, call B: tcgen05 FFT probe + host path re-test
mkdir -p gpurun_out
for v in 0 1 0x100 0x101; do
  timeout 60 python tools/tc_probe.py check $v > gpurun_out/r2b_tc_check_$v.json 2> gpurun_out/r2b_tc_check_$v.err; echo "check $v rc=$?"; cat gpurun_out/r2b_tc_check_$v.json | cut -c1-900; tail -2 gpurun_out/r2b_tc_check_$v.err
done
timeout 120 python tools/tc_probe.py time > gpurun_out/r2b_tc_time.json 2> gpurun_out/r2b_tc_time.err; echo "time rc=$?"; cat gpurun_out/r2b_tc_time.json; tail -2 gpurun_out/r2b_tc_time.err
timeout 600 python -m pytest tests/test_gpu_hostpath.py -x -q -m gpu > gpurun_out/r2b_hostpath.log 2>&1; echo "hostpath rc=$?"; tail -4 gpurun_out/r2b_hostpath.log
for t in 8 15; do for nt in 0 1; do SS_COPY_NT=$nt SS_COPY_THREADS=$t timeout 300 python bench.py --config c2 --steps 5 --no-cpu --e2e-modes pinned,pageable 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('copy_threads $t nt $nt pageable', d['e2e_pageable']['ms_per_step'], 'pinned', d['e2e']['ms_per_step'])"; done; done
