#!/bin/bash
# Round-1 evidence run (one B200): bench lines for every BASELINE config, the reference arm, the ncu
# launch list of the headline command and one `ncu --set full` capture per fused kernel.
# Everything lands in gpurun_out/; profiles/ gets the copies that are committed.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/r1_gpu.txt
for c in c2 c1 c3 c4 c5; do
  python bench.py --config $c > gpurun_out/bench_r1_$c.json 2> gpurun_out/bench_r1_$c.err
done
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r1_c2_reference.json 2>> gpurun_out/bench_r1_c2.err
SS_FAST512_GEN=2 python bench.py --config c2 --no-cpu --no-e2e > gpurun_out/bench_r1_c2_gen2.json 2>/dev/null
SS_STFT400_GEN=1 python bench.py --config c4 --no-cpu --no-e2e --steps 10 > gpurun_out/bench_r1_c4_gen1.json 2>/dev/null
SS_STFT2048_GEN=1 python bench.py --config c3 --no-cpu --no-e2e > gpurun_out/bench_r1_c3_gen1.json 2>/dev/null
python bench.py --config c2 --generic --no-cpu --no-e2e --steps 5 > gpurun_out/bench_r1_c2_generic.json 2>/dev/null
# launch list of the same command as the headline bench (short): per-launch durations, cold cache, serialised
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/launches_r1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_fast512c -c 1 -f -o gpurun_out/prof_fast512c_r1 python bench.py --config c2 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_r1_c2.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_fast512c -c 1 -f -o gpurun_out/prof_fast512c_c5_r1 python bench.py --config c5 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_r1_c5.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_stft400m -c 1 -f -o gpurun_out/prof_stft400m_r1 python bench.py --config c4 --clips 1024 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_r1_c4.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_stft2048m -c 1 -f -o gpurun_out/prof_stft2048m_r1 python bench.py --config c3 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_r1_c3.log 2>&1
for c in c2 c1 c3 c4 c5; do python -c "
import json
d=json.load(open('gpurun_out/bench_r1_$c.json')); print('$c', d['config']['kernel'], round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), 'e2e', d['e2e'] and round(d['e2e']['value']), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value']))"; done
