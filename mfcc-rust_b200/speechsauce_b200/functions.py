"""Mirror of speechsauce::functions' STFT (functions.rs:86-233)."""
from __future__ import annotations

import numpy as np

from . import _lib
from ._lib import check, lib


def stft2(signal, cfg, *, analysis_mem=None):
    """functions::stft2 (functions.rs:86-123): [ch, ttd] f32 -> [ch, T', F] complex64.
    The overlap memory lives in `cfg` and is carried across rows and calls like the reference's
    `SpeechConfig::analysis_mem`, unless an explicit `analysis_mem` array (updated in place) is given."""
    x = np.asarray(signal)
    if x.dtype != np.float32 or x.ndim != 2:
        raise TypeError("stft2 takes a 2-D float32 array")
    x = np.ascontiguousarray(x)
    ch, ttd = x.shape
    Tp = max(0, lib.ss_stft_num_frames(ttd, max(cfg.frame_size, 1)))
    out = np.empty((ch, Tp, cfg.freq_size), np.complex64)
    from .feature import _mem
    check(lib.ss_stft(cfg.plan, _lib.fptr(x), ch, ttd, ttd, _lib.fptr(out), _lib.fptr(_mem(analysis_mem, cfg))))
    return out


def stft1(signal, cfg, **kw):
    """functions::stft1 (functions.rs:199-233)."""
    x = np.asarray(signal)
    if x.ndim != 1:
        raise ValueError("stft1 takes a 1-D signal")
    return stft2(x[None, :], cfg, **kw)[0]
