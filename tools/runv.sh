# usage: bash tools/runv.sh <config> <variant ...>   -- times each lib variant ("default" = the in-tree build)
cfg=$1; shift
for v in "$@"; do
  if [ "$v" = default ]; then unset SPEECHSAUCE_CUDA_LIB; else export SPEECHSAUCE_CUDA_LIB=$PWD/mfcc-rust_b200/lib/variants/$v.so; fi
  python bench.py --config $cfg --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$cfg', '$v', round(d['ms_per_step'],3), round(d['roofline']['frac'],4))"
done
