#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2e_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r2e_gputests.log
for mb in 0 48; do for c in p_cmvn p_cmvn_long; do SS_CMVN_BATCH_MB=$mb timeout 120 python bench.py --config $c --steps 10 2>gpurun_out/r2e_$c.err | tee gpurun_out/r2e_${c}_$mb.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('batch_mb $mb $c', round(d['ms_per_step'],3),'ms', round(d['roofline']['achieved']),'GB/s frac', round(d['roofline']['frac'],3), 'launches', d['gpu_launches'])"; done; done
