#!/usr/bin/env python
"""PCIe ceiling of a box: concurrent pinned H2D + D2H cudaMemcpyAsync of given byte counts on every rank,
no kernels.  `python tools/pcie_ceiling.py [--h2d BYTES --d2h BYTES]` or under torchrun for N ranks; the
default sizes are the c2 step (4.19 GB in, 1.35 GB out).  bench.py runs the same measurement inline
(`pcie_ceiling` / `e2e.frac_of_pcie_ceiling`); this tool is for looking at a box on its own."""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--h2d", type=int, default=65536 * 16000 * 4)
ap.add_argument("--d2h", type=int, default=65536 * 97 * 53 * 4)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--streams", type=int, default=1, help="copies per direction are split over this many streams")
ap.add_argument("--register", action="store_true", help="also time cudaHostRegister / cudaHostUnregister of a pageable 1 GiB buffer (rank 0)")
a = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
hin = torch.empty(a.h2d // 4, dtype=torch.float32, pin_memory=True).fill_(1.0)
hout = torch.empty(a.d2h // 4, dtype=torch.float32, pin_memory=True)
din = torch.empty_like(hin, device=dev)
dout = torch.ones(a.d2h // 4, dtype=torch.float32, device=dev)
ss = [(torch.cuda.Stream(), torch.cuda.Stream()) for _ in range(a.streams)]


def once(h2d=True, d2h=True):
    for i, (si, so) in enumerate(ss):
        lo_i, hi_i = hin.numel() * i // a.streams, hin.numel() * (i + 1) // a.streams
        lo_o, hi_o = hout.numel() * i // a.streams, hout.numel() * (i + 1) // a.streams
        if h2d:
            with torch.cuda.stream(si):
                din[lo_i:hi_i].copy_(hin[lo_i:hi_i], non_blocking=True)
        if d2h:
            with torch.cuda.stream(so):
                hout[lo_o:hi_o].copy_(dout[lo_o:hi_o], non_blocking=True)
    torch.cuda.synchronize()


def timed(**kw):
    once(**kw)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(a.iters):
        once(**kw)
    dt = torch.tensor([(time.perf_counter() - t0) / a.iters], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return float(dt.item())


reg = None
if a.register and rank == 0:
    import numpy as np
    buf = np.ones(1 << 28, np.float32)  # 1 GiB, touched
    rt = torch.cuda.cudart()
    t0 = time.perf_counter()
    rc = rt.cudaHostRegister(buf.ctypes.data, buf.nbytes, 0)
    t1 = time.perf_counter()
    dbuf = torch.empty(1 << 28, dtype=torch.float32, device=dev)
    tb = torch.from_numpy(buf)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    dbuf.copy_(tb, non_blocking=True)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    rt.cudaHostUnregister(buf.ctypes.data)
    t4 = time.perf_counter()
    b2 = np.empty_like(buf)
    b2[...] = buf
    t5 = time.perf_counter()
    b2[...] = buf
    t6 = time.perf_counter()
    reg = {"register_rc": int(rc), "register_gb_s": buf.nbytes / (t1 - t0) / 1e9, "h2d_registered_gb_s": buf.nbytes / (t3 - t2) / 1e9,
           "unregister_gb_s": buf.nbytes / (t4 - t3) / 1e9, "numpy_copy_1_thread_gb_s": buf.nbytes / (t6 - t5) / 1e9}
both, up, down = timed(), timed(d2h=False), timed(h2d=False)
if rank == 0:
    print(json.dumps({"n_gpus": world, "h2d_bytes": a.h2d, "d2h_bytes": a.d2h, "streams_per_direction": a.streams,
                      "both_ms": both * 1e3, "h2d_only_ms": up * 1e3, "d2h_only_ms": down * 1e3,
                      "both_gb_s_per_gpu": (a.h2d + a.d2h) / both / 1e9, "both_gb_s_total": world * (a.h2d + a.d2h) / both / 1e9,
                      "h2d_gb_s_per_gpu": a.h2d / up / 1e9, "d2h_gb_s_per_gpu": a.d2h / down / 1e9, "host_register": reg}))
if world > 1:
    dist.destroy_process_group()
