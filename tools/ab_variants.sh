#!/bin/bash
# A/B of library variants built into gpurun_variants/*.so: parity subset + device-resident bench line per variant
# usage: tools/ab_variants.sh <config> [pytest -k expression]
cfg=${1:-c3}; kexpr=${2:-"stft or mel"}
for so in gpurun_variants/*.so; do
  export SPEECHSAUCE_CUDA_LIB=$PWD/$so
  t=$(timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$kexpr" 2>&1 | tail -1)
  for rep in 1 2; do
    timeout 200 python bench.py --config $cfg --steps 20 --warmup 5 --e2e-modes none 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$so', d['ms_per_step'], round(d['roofline']['frac'],4), '| $t')"
  done
done
