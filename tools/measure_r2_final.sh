#!/bin/bash
# Re-measures what changed after tools/measure_r2.sh ran (k_stft2048m's rotating n-tile pairs, coherent loads in the small
# cmvn kernels, two matrices per step in k_cmvn_coop) and re-runs the whole GPU test suite, the sanitizers and smoke() on
# the final binaries.  Output names match measure_r2.sh, so tools/collect_profiles_r2.py picks them up.
set -x
mkdir -p gpurun_out
python bench.py --config c3 --steps 10 > gpurun_out/r2z_bench_c3.json 2> gpurun_out/r2z_bench_c3.err
for c in p_cmvn p_cmvn_long; do
  python bench.py --config $c --steps 20 > gpurun_out/r2z_bench_$c.json 2> gpurun_out/r2z_bench_$c.err
done
ncu --set full --import-source on --clock-control none -k regex:k_stft2048m -c 1 -f -o gpurun_out/prof_stft2048m_r2 python bench.py --config c3 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_r2_c3.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_cmvn_reg -c 1 -f -o gpurun_out/prof_cmvn_reg_r2 python bench.py --config p_cmvn --steps 1 --warmup 1 > gpurun_out/ncu_r2_cmvn.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_cmvn_coop -c 1 -f -o gpurun_out/prof_cmvn_coop_r2 python bench.py --config p_cmvn_long --steps 1 --warmup 1 > gpurun_out/ncu_r2_cmvn_long.log 2>&1
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tests/sanitize_small.py > gpurun_out/r2z_sanitize_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2z_sanitize_$tool.log | tail -1)"
done
python -m pytest tests -q -m gpu > gpurun_out/r2z_gputests.log 2>&1; tail -3 gpurun_out/r2z_gputests.log
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/r2z_bench_c2.json 2> gpurun_out/r2z_bench_c2.err; tail -c 600 gpurun_out/r2z_bench_c2.json
