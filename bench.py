#!/usr/bin/env python
"""bench.py — throughput of the MFCC / mel hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config c2]

A "step" is one pass of the fused hot path over one batch of synthetic audio.  The default
workload is BASELINE.json configs[1] (c2): 65,536 x 1 s 16 kHz clips -> lmfe (40 filters) +
mfcc (13 cepstra) with per-clip CMVN.  `value` is whole-job audio-seconds per second with the
inputs resident in HBM; `e2e` is the same metric through the host-buffer C-ABI call
(`ss_features`: pinned host PCM -> H2D -> kernels -> D2H, all inside the timed region).
One process per GPU (torchrun); clips shard with no data-path collective.  `--scaling weak` (default):
every rank runs the named batch; `--scaling strong`: the named batch is split over the ranks (SURVEY 8e:
c2 8,192 clips per GPU at N = 8).  Beside `e2e` (pinned host buffers) the line carries `e2e_pageable`
(plain numpy arrays, what the reference's callers hand over), `e2e_inproc` (N > 1: ONE plan over all N
devices inside rank 0, the crate's `SpeechConfig::new` path) and `pcie_ceiling` (the same bytes as bare
concurrent H2D + D2H copies, no kernels).  `--impl reference` times the CPU restatement of the reference
(oracle/oracle_c.c, OpenMP over clips = "rayon over clips") on rank 0.  `--config p_*` times the
standalone public functions (processing::{preemphasis, stack_frames, cmvn, power_spectrum}).
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for p in (ROOT / "mfcc-rust_b200", ROOT / "oracle"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

import numpy as np  # noqa: E402

SPEECHPY = dict(sample_rate=16000, fft_points=512, frame_length=0.025, frame_stride=0.01, num_cepstral=13,
                num_filters=40, low_frequency=0.0, high_frequency=8000.0)
WORKLOADS = {
    # name: kind, config, clips per GPU, samples per clip, seconds per clip, outputs
    "c1": dict(kind="speechpy", cfg=SPEECHPY, n=1, length=16000, outs=("mfcc",), cmvn=0,
               desc="c1: mfcc on one 1 s 16 kHz clip (latency case)"),
    "c2": dict(kind="speechpy", cfg=SPEECHPY, n=65536, length=16000, outs=("lmfe", "mfcc"), cmvn=2,
               desc="c2: 65,536 x 1 s 16 kHz clips -> lmfe(40) + mfcc(13) with per-clip CMVN"),
    "c3": dict(kind="stft", cfg=dict(sample_rate=22050, fft_points=2048, frame_length=512 / 22050, frame_stride=0.01,
                                     num_cepstral=13, num_filters=128, low_frequency=0.0, high_frequency=11025.0),
               n=4096, length=110250, log=0, desc="c3: mel_spectrogram 22.05 kHz n_fft 2048 hop 512 128 mels, 4,096 x 5 s"),
    "c4": dict(kind="stft", cfg=dict(sample_rate=16000, fft_points=400, frame_length=0.01, frame_stride=0.01,
                                     num_cepstral=13, num_filters=80, low_frequency=0.0, high_frequency=8000.0),
               n=8192, length=480000, log=1, desc="c4: Whisper-shaped log-mel n_fft 400 hop 160 80 mels, 8,192 x 30 s"),
    "c5": dict(kind="speechpy", cfg=SPEECHPY, n=128, length=57_600_000, outs=("mfe", "energy", "ssc", "mfcc"), cmvn=0,
               desc="c5: 1-hour 16 kHz streams -> mfe + ssc + mfcc, 128 streams per GPU (1,024 over 8 GPUs)"),
}
# standalone public functions (processing.rs:31-53,65-129,179-181,265-300): HBM-bound streaming kernels
PRIMITIVES = {
    "p_preemphasis": "processing::preemphasis on 65,536 x 16,000 samples (shift 1, cof 0.98)",
    "p_stack": "processing::stack_frames of one 1-hour 16 kHz stream (L 400, S 160 -> 359,997 x 400)",
    "p_cmvn": "processing::cmvn(.., true) on 65,536 matrices of 97 x 13 (the c2 cepstra)",
    "p_cmvn_long": "processing::cmvn(.., true) on 64 matrices of 359,997 x 13 (1-hour cepstra)",
    "p_pspec": "processing::power_spectrum of 1,589,248 x 400 frames, n_fft 512",
}
FP32_PEAK_TFLOPS_AT_MAX_CLOCK = 148 * 128 * 2 * 1.965e9 / 1e12


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic(w, n_clips):
    """algorithmic bytes / flops per launch (DESIGN.md 'roofline'): PCM read once + outputs written once."""
    import speechsauce_oracle as o
    oc = o.SpeechConfig(**w["cfg"])
    if w["kind"] == "speechpy":
        L, S, N = oc.frame_sample_length, oc.frame_step, oc.window_size
        T = o.num_frames(w["length"], L, S, False)
        per = {"mfe": oc.num_filters, "energy": 1, "lmfe": oc.num_filters, "ssc": oc.num_filters, "mfcc": oc.num_cepstral}
        out_floats = T * sum(per[k] for k in w["outs"])
        taps = int((o.filterbanks(oc.num_filters, oc.freq_size, oc.sample_rate, oc.low_frequency, oc.high_frequency) != 0).sum())
        flops = T * (2.5 * N * np.log2(N) + 4 * oc.freq_size + 2 * taps + oc.freq_size + 2 * oc.num_filters + 2 * oc.num_cepstral * oc.num_filters)
    else:
        W, H = oc.window_size, oc.frame_size
        T = o.stft_num_frames(w["length"], H)
        out_floats = T * oc.num_filters
        taps = int((o.filterbanks(oc.num_filters, oc.freq_size, oc.sample_rate, oc.low_frequency, oc.high_frequency) != 0).sum())
        flops = T * (W + 2.5 * W * np.log2(W) + 4 * oc.freq_size + 2 * taps + oc.num_filters)
    bytes_per_clip = 4 * (w["length"] + out_floats)
    return dict(frames_per_clip=int(T), bytes=float(bytes_per_clip) * n_clips, flops=float(flops) * n_clips)


class ClockSampler:
    """samples SM clock + throttle reasons through NVML while the timed region runs"""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.005)

    def __enter__(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def bind_to_gpu_numa_node(index):
    """One process per GPU: run this process's host threads (and so first-touch its pinned buffers) on
    the CPUs NVML reports as local to the GPU, so eight concurrent H2D / D2H streams do not all pull
    from one socket's memory.  Returns the number of CPUs bound to, or None when NVML has no answer."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def bench_config(w, n, T):
    """the `config` object both arms print (identical when they time the same batch)"""
    return {"workload": w["desc"], "clips_per_gpu": n, "samples_per_clip": w["length"], "frames_per_clip": T,
            "l2": "inputs larger than L2" if n * w["length"] * 4 > (126 << 20) else "input fits in L2"}


def run_reference(args, w, n):
    """The reference arm: the CPU restatement (oracle port, C + OpenMP over clips) on rank 0.  It times the
    named batch itself where that is seconds of CPU work (c1, c2); the 30 s / 1 h shapes run a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_c
    import speechsauce_oracle as o
    # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host core it is allowed to
    cores = len(os.sched_getaffinity(0))
    cfg = oracle_c.make_config(**w["cfg"])
    if w["kind"] == "speechpy":
        sample = min(n, 65536) if w["length"] <= 10**6 else cores
        length = min(w["length"], 16000 * 600)
    else:
        sample = min(n, 8 * cores)
        length = w["length"]
    x = o.synth_signal(sample, length, 0x5EED0000 + int(args.config[1]))

    def step():
        if w["kind"] == "speechpy":
            kw = {k: True for k in w["outs"]}
            oracle_c.features(x, cfg, cmvn_mode=w["cmvn"], n_threads=cores, **kw)
        else:
            oracle_c.mel_spectrogram2(x, cfg, log=bool(w["log"]), n_threads=cores)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    secs = sample * length / w["cfg"]["sample_rate"]
    value = secs / dt
    alg = algorithmic(dict(w, length=length), sample)
    full = sample == n and length == w["length"]
    config = bench_config(w, n, algorithmic(w, 1)["frames_per_clip"])
    if not full:
        config["sample"] = f"{sample} clips x {length} samples per step (bounded sample of the named batch)"
    line = {
        "impl": "reference", "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s",
        "frames_per_s": alg["frames_per_clip"] * sample / dt,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config,
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} clips x {length / w['cfg']['sample_rate']:.0f} s per step"
                                   + (" (the named batch)" if full else " (bounded sample)") + "; restated CPU reference "
                                   "(oracle/oracle_c.c, C + OpenMP over clips), not the Rust binary (no Rust toolchain)"},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_primitive(args):
    """--config p_*: one standalone public function, device-resident, against the HBM roofline."""
    import torch
    import speechsauce_b200 as sb
    from speechsauce_b200 import _lib
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)
    name = args.config
    g = torch.Generator(device=dev).manual_seed(7)
    P = lambda t: C.c_void_p(t.data_ptr())
    if name == "p_preemphasis":
        x = torch.empty((65536, 16000), device=dev).normal_(0, 0.1, generator=g)
        y = torch.empty_like(x)
        step = lambda: _lib.check(sb.lib.ss_preemphasis_dev(sp, P(x), P(y), 65536, 16000, 16000, 1, 0.98))
        nbytes, units, kern = 2 * x.numel() * 4, 65536.0, "k_preemphasis"
    elif name == "p_stack":
        n = 57_600_000
        T = int(sb.lib.ss_num_frames(n, 400, 160, 0))
        x = torch.empty((n,), device=dev).normal_(0, 0.1, generator=g)
        y = torch.empty((T, 400), device=dev)
        step = lambda: _lib.check(sb.lib.ss_stack_frames_dev(sp, P(x), n, 400, 160, 0, 0, P(y), T))
        nbytes, units, kern = (n + T * 400) * 4, 3600.0, "k_stack_frames"
    elif name in ("p_cmvn", "p_cmvn_long"):
        nm, rows = (65536, 97) if name == "p_cmvn" else (64, 359_997)
        x = torch.empty((nm, rows, 13), device=dev).normal_(0, 1.0, generator=g)
        y = torch.empty_like(x)
        step = lambda: _lib.check(sb.lib.ss_cmvn_dev(sp, P(x), nm, rows, 13, 1, P(y)))
        nbytes, units, kern = 2 * x.numel() * 4, nm * rows / 100.0, "k_cmvn"
    else:
        rows = 1_589_248
        x = torch.empty((rows, 400), device=dev).normal_(0, 0.1, generator=g)
        y = torch.empty((rows, 257), device=dev)
        step = lambda: _lib.check(sb.lib.ss_power_spectrum_dev(sp, P(x), rows, 400, 512, P(y)))
        nbytes, units, kern = (x.numel() + y.numel()) * 4, rows / 100.0, "power_spectrum"
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    l0 = sb.lib.ss_kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(0) as clk:
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    peak, peak_src = peaks()
    ach = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps({
        "metric": "audio_seconds_per_second", "value": units / (ms * 1e-3), "unit": "audio-s/s", "n_gpus": 1,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": PRIMITIVES[name], "l2": "inputs larger than L2"}, "kernel": kern,
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": float(nbytes)},
        "cpu_baseline": None, "e2e": None, "gpu_launches": int(sb.lib.ss_kernel_launch_count() - l0),
        "clocks": clk.summary()}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=list(WORKLOADS) + list(PRIMITIVES))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--clips", type=int, default=0, help="override clips per GPU (debug)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-modes", default="pinned,pageable,inproc,ceiling",
                    help="which end-to-end legs to run (comma list of pinned, pageable, inproc, ceiling)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--generic", action="store_true", help="force the generic kernel (SS_FLAG_FORCE_GENERIC)")
    args = ap.parse_args()
    if args.config in PRIMITIVES:
        run_primitive(args)
        return
    w = dict(WORKLOADS[args.config])
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.clips:
        w["n"] = args.clips
    elif args.scaling == "strong" and args.config != "c5":  # c5 is 128 streams per GPU by definition (236 GB in total)
        w["n"] = max(1, w["n"] // max(world, args.gpus))
    if args.impl == "reference":
        run_reference(args, w, w["n"])
        return

    import torch
    import torch.distributed as dist
    import speechsauce_b200 as sb
    from speechsauce_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or sb.lib.ss_device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    cpu_pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        cpu_pg = dist.new_group(backend="gloo")  # host-side barrier for the in-process leg (no kernel spinning on the idle GPUs)

    cfg = sb.SpeechConfig(**w["cfg"], devices=[local], flags=_lib.SS_FLAG_FORCE_GENERIC if args.generic else 0)
    n, length = w["n"], w["length"]
    alg = algorithmic(w, n)
    T = alg["frames_per_clip"]
    gen = torch.Generator(device=dev)
    gen.manual_seed(0x5EED0000 + int(args.config[1]) + 1000 * rank)
    x = torch.empty((n, length), dtype=torch.float32, device=dev)
    for i in range(0, n, max(1, (1 << 28) // length)):  # generate in slabs to bound temporaries
        j = min(n, i + max(1, (1 << 28) // length))
        x[i:j].normal_(0.0, 0.1, generator=gen)
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)
    outs = {}
    KEYS = ("mfe", "energy", "lmfe", "ssc", "mfcc")

    def out_struct(ptr_of):
        return _lib.ss_feature_outputs(*[ptr_of(k) for k in KEYS], w.get("cmvn", 0), 0)

    if w["kind"] == "speechpy":
        dims = {"mfe": cfg.num_filters, "lmfe": cfg.num_filters, "ssc": cfg.num_filters, "mfcc": cfg.num_cepstral}
        for k in w["outs"]:
            outs[k] = torch.empty((n, T) if k == "energy" else (n, T, dims[k]), dtype=torch.float32, device=dev)
        o = out_struct(lambda k: C.c_void_p(outs[k].data_ptr()) if k in outs else None)

        def step():
            _lib.check(sb.lib.ss_features_dev(cfg.plan, 0, sp, C.c_void_p(x.data_ptr()), n, length, length, C.byref(o)))
    else:
        outs["mel"] = torch.empty((n, cfg.num_filters, T), dtype=torch.float32, device=dev)

        def step():
            _lib.check(sb.lib.ss_mel_spectrogram_dev(cfg.plan, 0, sp, C.c_void_p(x.data_ptr()), n, length, length,
                                                     C.c_void_p(outs["mel"].data_ptr()), None, None, w["log"]))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    l0 = sb.lib.ss_kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        barrier()
    launches = sb.lib.ss_kernel_launch_count() - l0
    ms_step = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    sr = w["cfg"]["sample_rate"]
    audio_s = n * length / sr
    value = world * audio_s / (ms_step * 1e-3)
    frames_s = world * n * T / (ms_step * 1e-3)

    # ---- end to end through the host-buffer C-ABI: H2D + kernels + D2H inside the timed region -------------
    e2e = e2e_pageable = e2e_inproc = ceiling = None
    modes = [] if args.no_e2e else [m for m in args.e2e_modes.split(",") if m]
    if modes:
        n_e = n
        while n_e * length * 4 > (6 << 30) and n_e > 1:  # keep the host buffers bounded (c4 / c5)
            n_e //= 2
        out_shapes = {k: (n_e,) + tuple(v.shape[1:]) for k, v in outs.items()}
        h2d = n_e * length * 4
        d2h = int(sum(int(np.prod(sh)) for sh in out_shapes.values()) * 4)
        ke = max(2, min(5, args.steps))

        def host_call(plan, n_rows, in_ptr, out_ptrs):
            if w["kind"] == "speechpy":
                ho = out_struct(lambda k: out_ptrs.get(k))
                _lib.check(sb.lib.ss_features(plan, in_ptr, n_rows, length, length, C.byref(ho)))
            else:
                _lib.check(sb.lib.ss_mel_spectrogram(plan, in_ptr, n_rows, length, length, out_ptrs["mel"], None, w["log"]))

        def timed(fn, ranks=world):
            fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(ke):
                fn()  # synchronous: returns after the last byte has reached the caller's buffer
            dt = max_over_ranks((time.perf_counter() - t0) / ke)
            return dt

        def leg(dt, units, **extra):
            d = {"value": units * n_e * length / sr / dt, "unit": "audio-s/s", "h2d_bytes_per_step": int(h2d),
                 "d2h_bytes_per_step": int(d2h), "clips_per_step": n_e, "ms_per_step": dt * 1e3, "steps": ke,
                 "gb_per_s_per_gpu": (h2d + d2h) / dt / 1e9}
            d.update(extra)
            return d

        key = "mfcc" if "mfcc" in outs else next(iter(outs))
        step()
        torch.cuda.synchronize()
        want = outs[key][1:5].cpu()  # row 0 skipped: the host STFT call carries the plan's overlap memory across calls
        try:
            hx = torch.empty((n_e, length), dtype=torch.float32, pin_memory=True)
            hx.copy_(x[:n_e])
            houts = {k: torch.empty(sh, dtype=torch.float32, pin_memory=True) for k, sh in out_shapes.items()}
            if "pinned" in modes:
                dt = timed(lambda: host_call(cfg.plan, n_e, C.c_void_p(hx.data_ptr()), {k: C.c_void_p(v.data_ptr()) for k, v in houts.items()}))
                e2e = leg(dt, world, host_cpus_bound=numa, host_memory="pinned (cudaHostAlloc)",
                          timing="host wall clock around the blocking ss_features / ss_mel_spectrogram call")
                assert torch.allclose(houts[key][1:5], want, atol=1e-5, rtol=1e-5), "e2e result differs from device result"
            if "ceiling" in modes:
                # the same bytes as bare copies: one H2D and one D2H stream per rank, all ranks at once, no kernels
                dbuf_in = x[:n_e]
                dbuf_out = {k: outs[k][:n_e] for k in outs}
                s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

                def copies():
                    with torch.cuda.stream(s_in):
                        dbuf_in.copy_(hx, non_blocking=True)
                    with torch.cuda.stream(s_out):
                        for k in houts:
                            houts[k].copy_(dbuf_out[k], non_blocking=True)
                    s_in.synchronize()
                    s_out.synchronize()
                dt = timed(copies)
                ceiling = leg(dt, world, what="concurrent pinned H2D + D2H cudaMemcpyAsync of the same bytes on every rank, no kernels")
                if e2e:
                    e2e["frac_of_pcie_ceiling"] = e2e["value"] / ceiling["value"]
            if "pageable" in modes:
                px = np.empty((n_e, length), np.float32)
                px[...] = hx.numpy()
                pouts = {k: np.empty(sh, np.float32) for k, sh in out_shapes.items()}
                dt = timed(lambda: host_call(cfg.plan, n_e, C.c_void_p(px.ctypes.data), {k: C.c_void_p(v.ctypes.data) for k, v in pouts.items()}))
                e2e_pageable = leg(dt, world, host_memory="pageable (np.empty)", copy_threads=int(sb.lib.ss_copy_threads()))
                assert np.allclose(pouts[key][1:5], want.numpy(), atol=1e-5, rtol=1e-5), "pageable e2e result differs from device result"
                if e2e:
                    e2e_pageable["vs_pinned"] = e2e_pageable["value"] / e2e["value"]
                if ceiling:
                    e2e_pageable["frac_of_pcie_ceiling"] = e2e_pageable["value"] / ceiling["value"]
                del px, pouts
            del hx, houts
            if "inproc" in modes and world > 1:
                # the crate's own multi-GPU path: ONE plan over all N devices in one process (rank 0), persistent
                # worker thread per device; the other ranks stay idle behind the barrier
                barrier()
                torch.cuda.synchronize()
                dist.barrier(group=cpu_pg)
                if rank == 0:
                    big = sb.SpeechConfig(**w["cfg"], devices=list(range(world)))
                    gx = torch.empty((world * n_e, length), dtype=torch.float32, pin_memory=True)
                    for r in range(world):
                        gx[r * n_e:(r + 1) * n_e].copy_(x[:n_e])
                    gouts = {k: torch.empty((world * n_e,) + sh[1:], dtype=torch.float32, pin_memory=True) for k, sh in out_shapes.items()}
                    call = lambda: host_call(big.plan, world * n_e, C.c_void_p(gx.data_ptr()), {k: C.c_void_p(v.data_ptr()) for k, v in gouts.items()})
                    call()
                    t0 = time.perf_counter()
                    for _ in range(ke):
                        call()
                    dt = (time.perf_counter() - t0) / ke
                    e2e_inproc = leg(dt, world, host_memory="pinned", what=f"one plan over {world} devices inside one process")
                    assert torch.allclose(gouts[key][1:5], want, atol=1e-5, rtol=1e-5), "in-process multi-device result differs"
                    if e2e:
                        e2e_inproc["vs_one_process_per_gpu"] = e2e_inproc["value"] / e2e["value"]
                    del gx, gouts
                    big.close()
                dist.barrier(group=cpu_pg)  # the other ranks wait here on the CPU: their GPUs belong to rank 0's plan meanwhile
                barrier()
        except RuntimeError as ex:  # pinned allocation failed
            if e2e is None:
                e2e = {"value": None, "error": str(ex)[:200]}

    cpu = None
    if rank == 0 and args.gpus == 1 and not args.no_cpu:
        import oracle_c
        import speechsauce_oracle as orc
        cores = len(os.sched_getaffinity(0))
        ccfg = oracle_c.make_config(**w["cfg"])
        if w["kind"] == "speechpy":
            cs = min(n, 2048 * cores) if length <= 10**6 else cores
            cl = min(length, 16000 * 600)
            xs = orc.synth_signal(cs, cl, 1)
            kw = {k: True for k in w["outs"]}
            cores = min(cores, cs)            # one clip (c1) is one thread's work: `cores` = the threads actually used
            reps = 200 if cs * cl <= 16000 * 64 else 1     # the latency case is a millisecond of work: repeat it
            oracle_c.features(xs[: max(1, cs // 8)], ccfg, cmvn_mode=w["cmvn"], n_threads=cores, **kw)
            t0 = time.perf_counter()
            for _ in range(reps):
                oracle_c.features(xs, ccfg, cmvn_mode=w["cmvn"], n_threads=cores, **kw)
        else:
            cs, cl, reps = min(n, 8 * cores), length, 1
            xs = orc.synth_signal(cs, cl, 1)
            t0 = time.perf_counter()
            oracle_c.mel_spectrogram2(xs, ccfg, log=bool(w["log"]), n_threads=cores)
        dt = (time.perf_counter() - t0) / reps
        cpu = {"value": cs * cl / sr / dt, "unit": "audio-s/s", "cores": cores, "kind": "port",
               "sample": f"{cs} clips x {cl / sr:.0f} s, {dt * reps:.1f} s wall"
                         + (f" over {reps} repetitions" if reps > 1 else "") + "; restated CPU reference "
                         "(oracle/oracle_c.c, C + OpenMP over clips), not the Rust binary"}

    if rank == 0:
        peak, peak_src = peaks()
        achieved = alg["bytes"] / (ms_step * 1e-3) / 1e9
        kernel = sb.lib.ss_last_kernel_name().decode()
        # dram__bytes_read.sum + dram__bytes_write.sum of the same kernel and workload from one
        # `ncu --set full` capture (profiles/ncu_traffic.json, per launch); null when there is none
        traffic, traffic_src = None, None
        tf = ROOT / "profiles" / "ncu_traffic.json"
        if tf.exists():
            t = json.loads(tf.read_text()).get(args.config)
            if t and t.get("clips") == n and t.get("kernel") == kernel:
                traffic, traffic_src = float(t["dram_bytes"]), t.get("source")
        line = {
            "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "frames_per_s": frames_s,
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(w, n, T), "kernel": kernel,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg["bytes"],
                         "fp32": {"achieved_tflops": alg["flops"] / (ms_step * 1e-3) / 1e12,
                                  "peak_tflops_at_max_clock": FP32_PEAK_TFLOPS_AT_MAX_CLOCK,
                                  "frac": alg["flops"] / (ms_step * 1e-3) / 1e12 / FP32_PEAK_TFLOPS_AT_MAX_CLOCK}},
            "cpu_baseline": cpu, "e2e": e2e, "e2e_pageable": e2e_pageable, "e2e_inproc": e2e_inproc,
            "pcie_ceiling": ceiling, "gpu_launches": int(launches), "clocks": clk.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
