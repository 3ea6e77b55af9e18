#!/usr/bin/env python
"""Numerics + rate probe of k_pspec512tc (tcgen05 FFT): python tc_probe.py selftest <0|1> | check <variant> | time"""
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
import subprocess  # noqa: E402

import torch  # noqa: E402

if not (HERE / "libtcfft.so").exists():
    subprocess.check_call(["make", "-C", str(HERE), "-s"])
_so = C.CDLL(str(HERE / "libtcfft.so"))
_vp, _i64 = C.c_void_p, C.c_int64
_so.tcfft_run.argtypes = [_vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _vp, C.c_int, _vp]
_so.tcfft_run.restype = C.c_int


class _Lib:
    @staticmethod
    def check(rc):
        if rc:
            raise RuntimeError(f"tcfft_run failed: {rc}")


class _Sb:
    class lib:
        ss_debug_pspec512tc = staticmethod(lambda *a: _so.tcfft_run(*a))


_lib, sb = _Lib, _Sb

P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
dev = torch.device("cuda", 0)


def run(x, n_clips, T, clip_stride, fs, L, variant, want_p=True, want_e=True, dbg=False):
    pspec = torch.zeros((n_clips * T, 257), device=dev) if want_p else None
    en = torch.zeros((n_clips * T,), device=dev) if want_e else None
    d = torch.zeros(32768, device=dev) if dbg else None
    s = torch.cuda.current_stream()
    _lib.check(sb.lib.ss_debug_pspec512tc(C.c_void_p(s.cuda_stream), P(x), n_clips, T, clip_stride, fs, L, x.numel(), P(pspec), P(en), variant, P(d)))
    torch.cuda.synchronize()
    return pspec, en, d


def check(variant):
    n_clips, clip_len, S, L = 3, 16000, 160, 400
    T = 97
    g = torch.Generator(device="cpu").manual_seed(3)
    xh = (torch.randn(n_clips, clip_len, generator=g) * 0.1)
    xh[1, 3000:3400] = torch.sin(torch.arange(400) * 0.3) * 0.5  # a tonal frame
    x = xh.to(dev)
    pspec, en, d = run(x, n_clips, T, clip_len, S, L, variant, dbg=True)
    xs = xh.numpy().astype(np.float64)
    fr = np.stack([[xs[c, t * S:t * S + L] for t in range(T)] for c in range(n_clips)]).reshape(n_clips * T, L)
    X = np.fft.rfft(fr, 512, axis=1)
    Pw = np.abs(X) / 512.0
    res = {"variant": variant}
    # stage 1 of tile 0: rows 16 j + n2 (frames 0..6 of clip 0)
    d1 = d[:4096].cpu().numpy().reshape(128, 32)
    y = np.zeros((7, 16, 32))
    for j in range(7):
        f = np.zeros(512)
        f[:L] = fr[j]
        m = f.reshape(32, 16)  # [n1][n2]
        Y = np.fft.fft(m, axis=0)  # [k1][n2]
        y[j, :, 0] = Y[0].real
        y[j, :, 1] = Y[16].real
        for k in range(1, 16):
            y[j, :, 2 * k] = Y[k].real
            y[j, :, 2 * k + 1] = Y[k].imag
    e1 = np.abs(d1[:112].reshape(7, 16, 32) - y)
    res["stage1_max_err"] = float(e1.max())
    res["stage1_scale"] = float(np.abs(y).max())
    res["stage1_err_by_frame"] = [float(v) for v in e1.max(axis=(1, 2))]
    # stage 2 of tile 0: rows 17 j + k1, cols (k2, re/im)
    d2 = d[4096:8192].cpu().numpy().reshape(128, 32)
    raw = d[8192:8192 + 92 * 16].cpu().numpy().reshape(92, 16)
    # un-swizzle (64B: 16-byte chunk c of row r sits at chunk c ^ ((r >> 1) & 3))
    uns = np.zeros_like(raw)
    for r in range(92):
        for c in range(4):
            uns[r, 4 * c:4 * c + 4] = raw[r, 4 * (c ^ ((r >> 1) & 3)):4 * (c ^ ((r >> 1) & 3)) + 4]
    want_raw = xs[0, :92 * 16].reshape(92, 16)
    res["tma_raw_nonzero"] = int((raw != 0).sum())
    res["tma_unswizzled_max_err"] = float(np.abs(uns - want_raw).max())
    res["tma_linear_max_err"] = float(np.abs(raw - want_raw).max())
    b1 = d[8192 + 4096:8192 + 4096 + 1024].cpu().numpy()
    res["b1_nonzero"] = int((b1 != 0).sum())
    res["b1_first"] = [float(v) for v in b1[:8]]
    a2 = d[8192 + 4096 + 1024:8192 + 4096 + 1024 + 4096].cpu().numpy()
    res["a2_nonzero"] = int((a2 != 0).sum())
    res["d1_nonzero"] = int((d[:4096] != 0).sum().item())
    res["d2_nonzero"] = int((d[4096:8192] != 0).sum().item())
    res["d1_row0"] = [float(v) for v in d[:8].cpu().numpy()]
    res["y_row0"] = [float(v) for v in y[0, 0, :8]]
    Xf = np.fft.fft(np.pad(fr[:7], ((0, 0), (0, 112))), axis=1) / 512.0
    want2 = np.zeros((119, 32))
    for j in range(7):
        for k1 in range(17):
            for k2 in range(16):
                want2[17 * j + k1, 2 * k2] = Xf[j, k1 + 32 * k2].real
                want2[17 * j + k1, 2 * k2 + 1] = Xf[j, k1 + 32 * k2].imag
    e2 = np.abs(d2[:119] - want2)
    res["stage2_max_err"] = float(e2.max())
    res["stage2_scale"] = float(np.abs(want2).max())
    got = pspec.cpu().numpy()
    err = np.abs(got - Pw)
    res["pspec_max_abs_err"] = float(err.max())
    res["pspec_max"] = float(Pw.max())
    res["pspec_max_rel_err_over_frame_max"] = float((err / Pw.max(axis=1, keepdims=True)).max())
    rel = err / np.maximum(Pw, 1e-30)
    res["pspec_rel_err_p999"] = float(np.quantile(rel, 0.999))
    res["pspec_rel_err_max"] = float(rel.max())
    res["energy_max_rel_err"] = float((np.abs(en.cpu().numpy() - Pw.sum(1)) / Pw.sum(1)).max())
    bad = np.argwhere(err > 1e-4 * Pw.max())
    res["n_bad"] = int(len(bad))
    res["first_bad"] = bad[:6].tolist()
    print(json.dumps(res))


def time_it():
    out = {}
    for name, n_clips, T, clip, fs, L in (("c2_framing", 65536, 97, 16000, 160, 400), ("pspec_api", 1, 1_589_248, 1_589_248 * 400, 400, 400)):
        x = torch.empty(n_clips * clip if n_clips > 1 else clip, device=dev).normal_(0, 0.1)
        for mode, wp, we in (("energy_only", False, True), ("pspec", True, False)):
            pspec = torch.zeros((n_clips * T, 257), device=dev) if wp else None
            en = torch.zeros((n_clips * T,), device=dev) if we else None
            s = torch.cuda.current_stream()
            call = lambda: _lib.check(sb.lib.ss_debug_pspec512tc(C.c_void_p(s.cuda_stream), P(x), n_clips, T, clip, fs, L, x.numel(), P(pspec), P(en), 0, None))
            for _ in range(3):
                call()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                call()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            out[f"{name}/{mode}"] = {"ms": ms, "frames_per_s": n_clips * T / ms * 1e3,
                                     "gb_s": ((x.numel() + (pspec.numel() if wp else 0) + (en.numel() if we else 0)) * 4) / ms / 1e6}
            del pspec, en
        del x
    print(json.dumps(out))


def selftest(bo):
    d = torch.zeros(32768, device=dev)
    s = torch.cuda.current_stream()
    _lib.check(sb.lib.ss_debug_pspec512tc(C.c_void_p(s.cuda_stream), None, 0, 0, 0, 0, 0, 0, None, None, 0x1000 | bo, P(d)))
    torch.cuda.synchronize()
    o = d[:16384].cpu().numpy().reshape(4, 128, 32)
    res = {"base_offset_mode": bo}
    want0 = np.zeros((128, 32))
    want0[:, :8] = np.arange(128)[:, None] * 100 + np.arange(8)[None, :]
    res["st_ld_ok"] = bool(np.array_equal(o[0][:, :8], want0[:, :8]))
    res["st_ld_row5"] = o[0][5, :10].tolist()
    m, k = np.meshgrid(np.arange(128), np.arange(32), indexing="ij")
    a12 = ((m * 32 + k) % 251).astype(np.float64)
    a3 = (((10 * (m >> 4) + k) * 16 + (m & 15)) % 251).astype(np.float64)
    for t, A in ((1, a12), (2, a12), (3, a3)):
        want = A * (k + 1)
        res[f"test{t}_exact"] = bool(np.array_equal(o[t], want))
        res[f"test{t}_nonzero"] = int((o[t] != 0).sum())
        res[f"test{t}_mismatch"] = int((o[t] != want).sum())
        res[f"test{t}_row17_got"] = o[t][17, :6].tolist()
        res[f"test{t}_row17_want"] = want[17, :6].tolist()
        # is the result a permutation-like match for another row?  report which logical row best matches physical row 17
        if not res[f"test{t}_exact"]:
            res[f"test{t}_rows_matching_any"] = int(sum(any(np.array_equal(o[t][r], want[q]) for q in range(128)) for r in range(128)))
    print(json.dumps(res))


if __name__ == "__main__":
    if sys.argv[1] == "selftest":
        selftest(int(sys.argv[2]))
        sys.exit(0)
    if sys.argv[1] == "check":
        check(int(sys.argv[2], 0))
    else:
        time_it()
