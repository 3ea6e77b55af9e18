#!/bin/bash
# 2 GPUs: the in-process multi-device plan (tests), torchrun bench with every e2e leg, the PCIe ceiling
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2f_topo.txt 2>&1; nproc >> gpurun_out/r2f_topo.txt; numactl -H >> gpurun_out/r2f_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_hostpath.py tests/test_gpu_parity.py -x -q -m gpu -k "multi_device or hostpath or current_device or two_streams or pinned or tiles or small_chunks" > gpurun_out/r2f_multigpu_tests.log 2>&1; echo "multi-gpu tests rc=$?"; tail -5 gpurun_out/r2f_multigpu_tests.log
for c in p_cmvn p_cmvn_long; do timeout 120 python bench.py --config $c --steps 10 2>/dev/null | tee gpurun_out/r2f_$c.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$c', round(d['ms_per_step'],3),'ms frac', round(d['roofline']['frac'],3))"; done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2f_bench_c2_2gpu.json 2> gpurun_out/r2f_bench_c2_2gpu.err; echo "bench 2gpu rc=$?"; tail -2 gpurun_out/r2f_bench_c2_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2f_bench_c2_2gpu.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','e2e_pageable','e2e_inproc','pcie_ceiling'):
    print(k, d.get(k))
PY
for n in 1 2; do timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 tools/pcie_ceiling.py 2>/dev/null | tail -1; done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/pcie_ceiling.py --streams 4 2>/dev/null | tail -1
