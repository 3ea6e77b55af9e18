#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
for v in base new; do
  if [ $v = base ]; then export SPEECHSAUCE_CUDA_LIB=$PWD/mfcc-rust_b200/lib/variants/base2048m.so; else unset SPEECHSAUCE_CUDA_LIB; fi
  timeout 200 python bench.py --config c3 --steps 30 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c3 $v rep $rep', round(d['ms_per_step'],4),'ms frac', round(d['roofline']['frac'],4))"
done; done
unset SPEECHSAUCE_CUDA_LIB
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stft or mel or cmvn" > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2i_tests.log
for c in p_cmvn p_cmvn_long; do timeout 120 python bench.py --config $c --steps 20 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$c', round(d['ms_per_step'],3),'ms frac', round(d['roofline']['frac'],3))"; done
