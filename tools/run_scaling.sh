#!/bin/bash
# usage: tools/run_scaling.sh N   (inside `gpurun --gpus N`): the north_star measurement matrix at N GPUs
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29600
run() {  # name, args...
  name=$1; shift
  port=$((port+1))
  timeout 600 $TR --master-port $port bench.py --gpus $N "$@" > gpurun_out/r2s_${name}_n$N.json 2> gpurun_out/r2s_${name}_n$N.err
  python - "$name" "$N" <<'PY'
import json,sys
name,n=sys.argv[1],sys.argv[2]
try:
    d=json.loads(open(f'gpurun_out/r2s_{name}_n{n}.json').read().strip().splitlines()[-1])
    g=lambda k:(d.get(k) or {}).get('value')
    print(name, 'N', n, 'ms', round(d['ms_per_step'],3), 'value %.3e'%d['value'], 'e2e', g('e2e'), 'pageable', g('e2e_pageable'), 'inproc', g('e2e_inproc'), 'ceiling', g('pcie_ceiling'))
except Exception as ex:
    print(name, 'N', n, 'FAILED', ex)
PY
}
nvidia-smi topo -m > gpurun_out/r2s_topo_n$N.txt 2>&1; nproc >> gpurun_out/r2s_topo_n$N.txt; free -g | head -2 >> gpurun_out/r2s_topo_n$N.txt
if [ "$N" = "8" ]; then
  timeout 600 python -m pytest tests/test_gpu_hostpath.py tests/test_gpu_parity.py -x -q -m gpu -k "multi_device or current_device" > gpurun_out/r2s_multigpu_tests_n$N.log 2>&1; echo "multi-gpu tests rc=$?"; tail -3 gpurun_out/r2s_multigpu_tests_n$N.log
fi
run c2_weak --config c2 --steps 10 --warmup 3 --no-cpu
run c2_strong --config c2 --scaling strong --steps 20 --warmup 3 --no-cpu --e2e-modes pinned,ceiling
if [ "$N" = "8" ]; then
  run c3_weak --config c3 --steps 10 --no-cpu --e2e-modes pinned,ceiling
  run c4_weak --config c4 --steps 5 --no-cpu --e2e-modes pinned,ceiling
  run c5_weak --config c5 --steps 5 --no-cpu --e2e-modes pinned,ceiling
  run c3_strong --config c3 --scaling strong --steps 20 --no-cpu --no-e2e
  run c4_strong --config c4 --scaling strong --steps 10 --no-cpu --no-e2e
else
  run c3_weak --config c3 --steps 10 --no-cpu --no-e2e
  run c4_weak --config c4 --steps 5 --no-cpu --no-e2e
  run c5_weak --config c5 --steps 5 --no-cpu --no-e2e
  run c3_strong --config c3 --scaling strong --steps 20 --no-cpu --no-e2e
  run c4_strong --config c4 --scaling strong --steps 10 --no-cpu --no-e2e
fi
port=$((port+1)); timeout 300 $TR --master-port $port tools/pcie_ceiling.py 2>/dev/null | tail -1 | tee gpurun_out/r2s_pcie_n$N.json
