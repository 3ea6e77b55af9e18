# ncu captures of the two TMA-staged kernels (one launch each)
ncu --set full --import-source on --clock-control none -k regex:k_fast512c -c 1 -f -o gpurun_out/prof_fast512c_r1t python bench.py --config c2 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_ct.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_stft400m -c 1 -f -o gpurun_out/prof_stft400m_r1t python bench.py --config c4 --clips 1024 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_400t.log 2>&1
