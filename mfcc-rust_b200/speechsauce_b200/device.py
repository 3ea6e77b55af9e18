"""Device-resident calls: inputs are anything that speaks `__cuda_array_interface__` (torch tensors, cupy arrays,
numba device arrays), outputs are device arrays that speak it too -- no host round trip, no copy.

This is what the py-speechsauce binding's `to_pyarray` copy (py-speechsauce/src/lib.rs:176,203) turns into when
the caller's data already lives on the GPU.  Work is enqueued on `stream` (an int handle, an
object with `.cuda_stream`, or None = the legacy default stream) and NOT synchronised; outputs come from
`allocator(shape) -> array with __cuda_array_interface__` (default: torch.empty on the input's device)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, lib


class _View:
    """pointer / shape of a C-contiguous float32 device array"""

    def __init__(self, x, ndim):
        cai = getattr(x, "__cuda_array_interface__", None)
        if cai is None:
            raise TypeError("expected a device array (an object with __cuda_array_interface__)")
        if np.dtype(cai["typestr"]) != np.float32:
            raise TypeError("speechsauce is f32-only (README.md:17); got %s" % cai["typestr"])
        shape = tuple(int(s) for s in cai["shape"])
        if len(shape) == ndim - 1:
            shape = (1,) + shape
            self.squeeze = True
        else:
            self.squeeze = False
        if len(shape) != ndim:
            raise ValueError(f"expected a {ndim - 1}-D or {ndim}-D array")
        strides = cai.get("strides")
        if strides is not None:
            want, acc = [], 4
            for s in reversed(tuple(int(s) for s in cai["shape"])):
                want.append(acc)
                acc *= s
            if tuple(int(s) for s in strides) != tuple(reversed(want)):
                raise ValueError("device arrays must be C-contiguous")
        self.ptr = C.c_void_p(int(cai["data"][0]))
        self.shape = shape
        self.obj = x


def _device_of(x):
    d = getattr(x, "device", None)
    idx = getattr(d, "index", None)
    if idx is None:
        idx = getattr(d, "id", None)  # cupy
    return 0 if idx is None else int(idx)


def _stream(stream):
    if stream is None:
        return None
    return C.c_void_p(int(getattr(stream, "cuda_stream", stream)))


def _dev_index(cfg, device):
    devs = cfg.devices
    if devs is None:
        return device
    if device not in devs:
        raise ValueError(f"the config has no plan on cuda:{device} (devices={devs})")
    return devs.index(device)


def _torch_alloc(device):
    import torch
    return lambda shape: torch.empty(shape, dtype=torch.float32, device=torch.device("cuda", device))


def _ptr(a):
    return None if a is None else C.c_void_p(int(a.__cuda_array_interface__["data"][0]))


def features(signal, cfg, *, mfe=False, energy=False, lmfe=False, ssc=False, mfcc=False, cmvn_mode=0, stream=None,
             allocator=None):
    """ss_features_dev: any subset of the feature outputs from ONE pass over device-resident PCM [n, clip_len]."""
    v = _View(signal, 2)
    n, clip_len = v.shape
    device = _device_of(signal)
    alloc = allocator or _torch_alloc(device)
    T = max(0, int(lib.ss_num_frames(clip_len, cfg.frame_sample_length, cfg.frame_step, 0)))
    nf, nc = cfg.num_filters, cfg.num_cepstral
    out = {}
    if mfe: out["mfe"] = alloc((n, T, nf))
    if energy: out["energy"] = alloc((n, T))
    if lmfe: out["lmfe"] = alloc((n, T, nf))
    if ssc: out["ssc"] = alloc((n, T, nf))
    if mfcc: out["mfcc"] = alloc((n, T, nc))
    o = _lib.ss_feature_outputs(*[_ptr(out.get(k)) for k in ("mfe", "energy", "lmfe", "ssc", "mfcc")], int(cmvn_mode), 0)
    check(lib.ss_features_dev(cfg.plan, _dev_index(cfg, device), _stream(stream), v.ptr, n, clip_len, clip_len, C.byref(o)))
    return {k: (a[0] if v.squeeze else a) for k, a in out.items()}


def mfcc(signal, cfg, cmvn_mode=0, **kw):
    """feature::mfcc (feature.rs:99-148) on the device."""
    return features(signal, cfg, mfcc=True, cmvn_mode=cmvn_mode, **kw)["mfcc"]


def mel_spectrogram2(signal, cfg, *, log=False, stream=None, allocator=None):
    """feature::mel_spectrogram2 (feature.rs:163-174) on the device: [ch, ttd] -> [ch, nf, T'].  Starts from a zero
    overlap memory (a fresh SpeechConfig, config.rs:162); the plan's host-side memory is not touched."""
    v = _View(signal, 2)
    ch, ttd = v.shape
    device = _device_of(signal)
    alloc = allocator or _torch_alloc(device)
    Tp = max(0, int(lib.ss_stft_num_frames(ttd, max(cfg.frame_size, 1))))
    out = alloc((ch, cfg.num_filters, Tp))
    check(lib.ss_mel_spectrogram_dev(cfg.plan, _dev_index(cfg, device), _stream(stream), v.ptr, ch, ttd, ttd, _ptr(out), None, None,
                                     int(log)))
    return out[0] if v.squeeze else out


def stft2(signal, cfg, *, stream=None, allocator=None):
    """functions::stft2 (functions.rs:86-123) on the device: [ch, ttd] -> [ch, T', F, 2] (re, im interleaved)."""
    v = _View(signal, 2)
    ch, ttd = v.shape
    device = _device_of(signal)
    alloc = allocator or _torch_alloc(device)
    Tp = max(0, int(lib.ss_stft_num_frames(ttd, max(cfg.frame_size, 1))))
    out = alloc((ch, Tp, cfg.freq_size, 2))
    check(lib.ss_stft_dev(cfg.plan, _dev_index(cfg, device), _stream(stream), v.ptr, ch, ttd, ttd, _ptr(out), None, None))
    return out[0] if v.squeeze else out


def cmvn(vec, variance_normalization=False, *, stream=None, allocator=None):
    """processing::cmvn (processing.rs:265-300) on the device: [rows, cols] or a batch [n, rows, cols]."""
    v = _View(vec, 3)
    n, rows, cols = v.shape
    alloc = allocator or _torch_alloc(_device_of(vec))
    out = alloc((n, rows, cols))
    check(lib.ss_cmvn_dev(_stream(stream), v.ptr, n, rows, cols, int(bool(variance_normalization)), _ptr(out)))
    return out[0] if v.squeeze else out


def preemphasis(signal, shift=1, cof=0.98, *, stream=None, allocator=None):
    """processing::preemphasis (processing.rs:31-53) on the device: [n] or a batch [rows, n]."""
    v = _View(signal, 2)
    rows, n = v.shape
    alloc = allocator or _torch_alloc(_device_of(signal))
    out = alloc((rows, n))
    check(lib.ss_preemphasis_dev(_stream(stream), v.ptr, _ptr(out), rows, n, n, int(shift), float(cof)))
    return out[0] if v.squeeze else out
