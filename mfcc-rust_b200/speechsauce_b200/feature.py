"""Mirror of speechsauce::feature (feature.rs): mfcc, mfe, lmfe, ssc, mel_spectrogram1/2.

1-D inputs follow the reference signatures; 2-D inputs ([n_clips, clip_len]) are the batched
form the GPU path is built for (clips are sharded over the config's devices)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, lib


def _rows(signal):
    x = np.asarray(signal)
    if x.dtype != np.float32:
        raise TypeError("speechsauce is f32-only (README.md:17); got %s" % x.dtype)
    one = x.ndim == 1
    if one:
        x = x[None, :]
    if x.ndim != 2:
        raise ValueError("signal must be 1-D or 2-D")
    return np.ascontiguousarray(x), one  # as_standard_layout(), SURVEY 8(b)


def _mem(analysis_mem, cfg):
    """the caller-held STFT overlap memory (config.rs:126): exactly W - H float32, C-contiguous; the library reads
    and writes that many floats through the pointer, so anything else is refused here"""
    if analysis_mem is None:
        return None
    m = analysis_mem
    want = cfg.window_size - cfg.frame_size
    if not isinstance(m, np.ndarray) or m.dtype != np.float32 or not m.flags.c_contiguous or not m.flags.writeable or m.size != want:
        raise ValueError(f"analysis_mem must be a writable C-contiguous float32 array of window_size - frame_size = {want} elements")
    return m


def _frames(cfg, clip_len):
    T = lib.ss_num_frames(clip_len, cfg.frame_sample_length, cfg.frame_step, 0)
    return T


def features(signal, cfg, *, mfe=False, energy=False, lmfe=False, ssc=False, mfcc=False, cmvn_mode=0, empty=np.empty):
    """ss_features: any subset of outputs from ONE pass over the PCM.  `empty`: the output allocator
    (np.empty; hostmem.pinned_empty gives page-locked results that are DMA'd in place)."""
    x, one = _rows(signal)
    n, clip_len = x.shape
    T = _frames(cfg, clip_len)
    if T < 1:  # let the library produce the error message
        T = 0
    nf, nc = cfg.num_filters, cfg.num_cepstral
    out = {}
    if mfe: out["mfe"] = empty((n, T, nf), np.float32)
    if energy: out["energy"] = empty((n, T), np.float32)
    if lmfe: out["lmfe"] = empty((n, T, nf), np.float32)
    if ssc: out["ssc"] = empty((n, T, nf), np.float32)
    if mfcc: out["mfcc"] = empty((n, T, nc), np.float32)
    o = _lib.ss_feature_outputs(*[_lib.fptr(out.get(k)) for k in ("mfe", "energy", "lmfe", "ssc", "mfcc")],
                                int(cmvn_mode), 0)
    check(lib.ss_features(cfg.plan, _lib.fptr(x), n, clip_len, clip_len, C.byref(o)))
    return {k: (v[0] if one else v) for k, v in out.items()}


def librosa_config(sr, n_fft=2048, hop_length=512, n_mels=128, fmin=0.0, fmax=None, **kw):
    """A plan for `melspectrogram_librosa` (SS_FLAG_LIBROSA).  frame_length carries the hop: the library
    derives hop = trunc(frame_length * sample_rate) in f32 like config.rs:154, hence the half-sample margin."""
    from .config import SpeechConfig
    flags = _lib.SS_FLAG_LIBROSA | kw.pop("flags", 0)
    return SpeechConfig(sample_rate=sr, fft_points=n_fft, frame_length=(hop_length + 0.5) / sr, frame_stride=hop_length / sr,
                        num_cepstral=min(13, n_mels), num_filters=n_mels, low_frequency=fmin, high_frequency=fmax,
                        flags=flags, **kw)


def melspectrogram_librosa(y, cfg, db=False):
    """librosa.feature.melspectrogram(y, sr, n_fft, hop_length, window="hann", center=True,
    pad_mode="reflect", power=2.0, n_mels, fmin, fmax, htk=False, norm="slaney") -> [..., n_mels, 1 + n / hop];
    `db`: librosa.power_to_db(S, ref=1.0, amin=1e-10, top_db=None)."""
    x, one = _rows(y)
    n, ttd = x.shape
    T = int(lib.ss_librosa_num_frames(ttd, cfg.frame_size))
    out = np.empty((n, cfg.num_filters, max(T, 0)), np.float32)
    check(lib.ss_librosa_melspectrogram(cfg.plan, _lib.fptr(x), n, ttd, ttd, _lib.fptr(out), int(bool(db))))
    return out[0] if one else out


def mfe(signal, cfg):
    """feature::mfe (feature.rs:200-233) -> (features [T, nf], frame_energies [T])."""
    r = features(signal, cfg, mfe=True, energy=True)
    return r["mfe"], r["energy"]


def lmfe(signal, cfg):
    """feature::lmfe (feature.rs:242-245)."""
    return features(signal, cfg, lmfe=True)["lmfe"]


def ssc(signal, cfg):
    """feature::ssc — named by README.md:15, defined in DESIGN.md."""
    return features(signal, cfg, ssc=True)["ssc"]


def mfcc(signal, cfg, cmvn_mode=0):
    """feature::mfcc (feature.rs:99-148); cmvn_mode applies processing::cmvn per clip."""
    return features(signal, cfg, mfcc=True, cmvn_mode=cmvn_mode)["mfcc"]


def mel_spectrogram2(signal, cfg, *, log=False, analysis_mem=None):
    """feature::mel_spectrogram2 (feature.rs:163-174): [ch, ttd] -> [ch, nf, T']."""
    x, _ = _rows(signal)
    ch, ttd = x.shape
    Tp = max(0, lib.ss_stft_num_frames(ttd, max(cfg.frame_size, 1)))
    out = np.empty((ch, cfg.num_filters, Tp), np.float32)
    check(lib.ss_mel_spectrogram(cfg.plan, _lib.fptr(x), ch, ttd, ttd, _lib.fptr(out),
                                 _lib.fptr(_mem(analysis_mem, cfg)), int(log)))
    return out


def mel_spectrogram1(signal, cfg, **kw):
    """feature::mel_spectrogram1 (feature.rs:151-162), intended semantics: one-row mel_spectrogram2."""
    x = np.asarray(signal)
    if x.ndim != 1:
        raise ValueError("mel_spectrogram1 takes a 1-D signal")
    return mel_spectrogram2(x[None, :], cfg, **kw)[0]


def extract_derivative_feature(feat):
    """feature::extract_derivative_feature (feature.rs:253-269, private in the reference):
    [T, C] -> [T, C, 3] = (static, delta, delta-delta)."""
    x = np.ascontiguousarray(feat, np.float32)
    if x.ndim != 2:
        raise ValueError("feature must be 2-D")
    out = np.empty(x.shape + (3,), np.float32)
    check(lib.ss_extract_derivative_feature(_lib.fptr(x), x.shape[0], x.shape[1], _lib.fptr(out)))
    return out
