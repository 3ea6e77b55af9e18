"""Mirror of speechsauce::processing (processing.rs): preemphasis, stack_frames, power_spectrum, cmvn."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, lib


def _f32(a, ndim=None):
    a = np.asarray(a)
    if a.dtype != np.float32:
        raise TypeError("speechsauce is f32-only (README.md:17); got %s" % a.dtype)
    if ndim is not None and a.ndim != ndim:
        raise ValueError(f"expected a {ndim}-D array")
    return np.ascontiguousarray(a)


def preemphasis(signal, shift=1, cof=0.98):
    """processing::preemphasis (processing.rs:31-53); 1-D, or 2-D = one signal per row."""
    x = _f32(signal)
    out = np.empty_like(x)
    if x.ndim == 1:
        check(lib.ss_preemphasis(_lib.fptr(x), _lib.fptr(out), x.shape[0], int(shift), float(cof)))
    elif x.ndim == 2:
        check(lib.ss_preemphasis_batch(_lib.fptr(x), _lib.fptr(out), x.shape[0], x.shape[1], x.shape[1],
                                       int(shift), float(cof)))
    else:
        raise ValueError("signal must be 1-D or 2-D")
    return out


def stack_frames(signal, sample_rate, frame_length=0.020, frame_stride=0.020, filter=None,
                 zero_padding=True, *, literal=False):
    """processing::stack_frames (processing.rs:65-129).  `filter` must be None (the reference's
    filter path has a shape bug and is never used by the features, Q4)."""
    if filter is not None:
        raise NotImplementedError("stack_frames(filter=Some(..)) is broken in the reference (processing.rs:122-126)")
    x = _f32(signal, 1)
    sr = np.float32(sample_rate)
    L = int(np.floor(abs(float(sr * np.float32(frame_length))) + 0.5))   # f32::round
    S = int(np.floor(abs(float(sr * np.float32(frame_stride))) + 0.5))
    T = C.c_int64(0)
    flags = _lib.SS_FLAG_FRAMING_LITERAL if literal else 0
    check(lib.ss_stack_frames(_lib.fptr(x), x.shape[0], L, S, int(zero_padding), flags, None, C.byref(T)))
    out = np.empty((T.value, L), np.float32)
    check(lib.ss_stack_frames(_lib.fptr(x), x.shape[0], L, S, int(zero_padding), flags, _lib.fptr(out), C.byref(T)))
    return out


def power_spectrum(frames, fft_points=512):
    """processing::power_spectrum (processing.rs:179-181): (1/fft_points) * |rfft(frame)|."""
    f = _f32(frames, 2)
    out = np.empty((f.shape[0], fft_points // 2 + 1), np.float32)
    check(lib.ss_power_spectrum(_lib.fptr(f), f.shape[0], f.shape[1], int(fft_points), _lib.fptr(out)))
    return out


def log_power_spectrum(frames, fft_points=512, normalize=True):
    """processing::log_power_spectrum (processing.rs:194-214, private in the reference): 10 log10 of the
    power spectrum (-200 below 1e-20); `normalize` subtracts the last value above the threshold."""
    f = _f32(frames, 2)
    out = np.empty((f.shape[0], fft_points // 2 + 1), np.float32)
    check(lib.ss_log_power_spectrum(_lib.fptr(f), f.shape[0], f.shape[1], int(fft_points), int(bool(normalize)), _lib.fptr(out)))
    return out


def cmvn(vec, variance_normalization=False):
    """processing::cmvn (processing.rs:265-300); 2-D, or 3-D = a batch of matrices."""
    v = _f32(vec)
    out = np.empty_like(v)
    if v.ndim == 2:
        check(lib.ss_cmvn(_lib.fptr(v), v.shape[0], v.shape[1], int(bool(variance_normalization)), _lib.fptr(out)))
    elif v.ndim == 3:
        check(lib.ss_cmvn_batch(_lib.fptr(v), v.shape[0], v.shape[1], v.shape[2],
                                int(bool(variance_normalization)), _lib.fptr(out)))
    else:
        raise ValueError("vec must be 2-D or 3-D")
    return out


def derivative_extraction(feat, delta_windows=2):
    """processing::derivative_extraction (processing.rs:222-254), numpy 'edge' padding along features."""
    f = _f32(feat, 2)
    out = np.empty_like(f)
    check(lib.ss_derivative_extraction(_lib.fptr(f), f.shape[0], f.shape[1], int(delta_windows), _lib.fptr(out)))
    return out


def cmvnw(vec, win_size=301, variance_normalization=False):
    """processing::cmvnw (processing.rs:315-371), numpy 'symmetric' padding along time."""
    v = _f32(vec, 2)
    out = np.empty_like(v)
    check(lib.ss_cmvnw(_lib.fptr(v), v.shape[0], v.shape[1], int(win_size), int(bool(variance_normalization)), _lib.fptr(out)))
    return out
