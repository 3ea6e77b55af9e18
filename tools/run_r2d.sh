#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/tc_probe.py time > gpurun_out/r2d_tc_time.json 2> gpurun_out/r2d_tc_time.err; echo "time rc=$?"; cat gpurun_out/r2d_tc_time.json
timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_uniform.sum --clock-control none -k regex:k_pspec512tc -s 3 -c 1 --csv --log-file gpurun_out/r2d_tc_ncu.csv python tools/tc_probe.py time > /dev/null 2>&1; tail -6 gpurun_out/r2d_tc_ncu.csv
timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_fast512c -s 3 -c 1 --csv --log-file gpurun_out/r2d_f512c_ncu.csv python bench.py --config c2 --steps 2 --no-cpu --no-e2e > /dev/null 2>&1; tail -4 gpurun_out/r2d_f512c_ncu.csv
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2d_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r2d_gputests.log
for c in p_preemphasis p_stack p_cmvn p_cmvn_long p_pspec; do SS_PSPEC_TC=0 timeout 120 python bench.py --config $c --steps 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$c', round(d['ms_per_step'],3),'ms', round(d['roofline']['achieved']),'GB/s frac', round(d['roofline']['frac'],3))"; done
