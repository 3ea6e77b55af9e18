#!/bin/bash
# usage: tools/run_scaling_c34.sh N   (inside `gpurun --gpus N`): device-resident c3 / c4 lines at N GPUs, re-measured after
# the kernels changed late in round 2 (the e2e legs are PCIe-bound and keep the numbers of tools/run_scaling.sh)
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29700
run() {  # name, args...
  name=$1; shift
  port=$((port+1))
  timeout 300 $TR --master-port $port bench.py --gpus $N "$@" > gpurun_out/r2t_${name}_n$N.json 2> gpurun_out/r2t_${name}_n$N.err
  python - "$name" "$N" <<'PY'
import json,sys
name,n=sys.argv[1],sys.argv[2]
try:
    d=json.loads(open(f'gpurun_out/r2t_{name}_n{n}.json').read().strip().splitlines()[-1])
    print(name, 'N', n, 'ms', round(d['ms_per_step'],3), 'value %.3e'%d['value'])
except Exception as ex:
    print(name, 'N', n, 'FAILED', ex)
PY
}
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_hostpath.py tests/test_gpu_parity.py -x -q -m gpu -k "multi_device or current_device or tiles or cmvn" > gpurun_out/r2t_multigpu_tests_n$N.log 2>&1; echo "multi-gpu tests rc=$?"; tail -3 gpurun_out/r2t_multigpu_tests_n$N.log
fi
run c3_weak --config c3 --steps 20 --no-cpu --no-e2e
run c4_weak --config c4 --steps 10 --no-cpu --no-e2e
run c3_strong --config c3 --scaling strong --steps 20 --no-cpu --no-e2e
run c4_strong --config c4 --scaling strong --steps 10 --no-cpu --no-e2e
