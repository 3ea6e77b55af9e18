#!/bin/bash
# Round-2 evidence run (one B200, final binaries): bench lines for every BASELINE config and the standalone functions,
# the reference arm, the ncu launch list of the headline command, one `ncu --set full` capture per kernel, sanitizers.
# Everything lands in gpurun_out/r2z_*; tools/collect_profiles_r2.py turns it into the committed files under profiles/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/r2z_gpu.txt; nproc >> gpurun_out/r2z_gpu.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r2z_bench_c2.json 2> gpurun_out/r2z_bench_c2.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2z_bench_c2_reference.json 2>> gpurun_out/r2z_bench_c2.err
for c in c1 c3 c4 c5; do
  python bench.py --config $c --steps 10 > gpurun_out/r2z_bench_$c.json 2> gpurun_out/r2z_bench_$c.err
done
for c in p_preemphasis p_stack p_cmvn p_cmvn_long p_pspec; do
  python bench.py --config $c --steps 20 > gpurun_out/r2z_bench_$c.json 2> gpurun_out/r2z_bench_$c.err
done
python bench.py --config c2 --generic --no-cpu --no-e2e --steps 5 > gpurun_out/r2z_bench_c2_generic.json 2>/dev/null
SS_PSPEC_GENERIC=1 python bench.py --config p_pspec --steps 5 > gpurun_out/r2z_bench_p_pspec_generic.json 2>/dev/null
# launch list of the same command as the headline bench (short): per-launch durations, cold cache, serialised
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 5 --warmup 3 --no-cpu --e2e-modes pinned > gpurun_out/launches_r2.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_fast512c -c 1 -f -o gpurun_out/prof_fast512c_r2 python bench.py --config c2 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_r2_c2.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_fast512c -c 1 -f -o gpurun_out/prof_fast512c_c5_r2 python bench.py --config c5 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_r2_c5.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_stft400m -c 1 -f -o gpurun_out/prof_stft400m_r2 python bench.py --config c4 --clips 1024 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_r2_c4.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_stft2048m -c 1 -f -o gpurun_out/prof_stft2048m_r2 python bench.py --config c3 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_r2_c3.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_pspec512 -c 1 -f -o gpurun_out/prof_pspec512_r2 python bench.py --config p_pspec --steps 1 --warmup 1 > gpurun_out/ncu_r2_pspec.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_cmvn_reg -c 1 -f -o gpurun_out/prof_cmvn_reg_r2 python bench.py --config p_cmvn --steps 1 --warmup 1 > gpurun_out/ncu_r2_cmvn.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_cmvn_coop -c 1 -f -o gpurun_out/prof_cmvn_coop_r2 python bench.py --config p_cmvn_long --steps 1 --warmup 1 > gpurun_out/ncu_r2_cmvn_long.log 2>&1
SS_CMVN_COOP=0 python bench.py --config p_cmvn_long --steps 20 > gpurun_out/r2z_bench_p_cmvn_long_3pass.json 2>/dev/null
SS_CMVN_REG=0 python bench.py --config p_cmvn --steps 20 > gpurun_out/r2z_bench_p_cmvn_smem.json 2>/dev/null
python tools/time_stft.py > gpurun_out/r2z_time_stft.txt 2>&1
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tests/sanitize_small.py > gpurun_out/r2z_sanitize_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2z_sanitize_$tool.log | tail -1)"
done
python -m pytest tests -q -m gpu > gpurun_out/r2z_gputests.log 2>&1; tail -3 gpurun_out/r2z_gputests.log
python __graft_entry__.py smoke 2>&1 | tail -2
