"""CPU restatement ("oracle") of speechsauce's MFCC / mel-spectrogram hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (`mfcc-rust_b200/`) may import
this module; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU baseline do.

PARITY UNPINNED: the reference (secretsauceai/mfcc-rust) is a Rust crate; there is no
Rust toolchain in this image, its numeric dependencies are not vendored, and its own
tests hold no golden vectors (they check shapes and "no NaN" only,
`speechsauce/src/lib.rs:41-134`).  This file therefore restates the reference's
algorithm line by line and is pinned by
  * the shapes / invariants the reference's tests assert (lib.rs:41-134),
  * closed-form known-answer tests (tests/test_oracle_kat.py), and
  * the published conventions of the third-party crates that carry the arithmetic:
      ndrustfft 0.4 (`ndfft_r2c`: unnormalised forward FFT, e^{-2 pi i kn/N};
                     `nddct2`: 2 * sum x[n] cos(pi k (2n+1) / 2N), i.e. scipy's dct type 2),
      realfft 3.2   (unnormalised forward R2C FFT of any length),
      ndarray 0.15  (`linspace` = start + step*i, sequential row accumulation for
                     `mean_axis(0)`, Welford for `std_axis`),
      einsum-derive 0.1 (dense f32 contraction).
    The one convention that cannot be verified here is the DCT-II scale; it is kept as
    the single named constant `DCT2_SCALE`.

Every function cites the reference file:line (relative to /root/reference/) it follows.
`dtype=np.float32` mirrors the reference's arithmetic ("hardcoded to use f32",
README.md:17); `dtype=np.float64` is the same algorithm in double precision and is used
only to bound rounding error.  Index computations (frame counts, filterbank bin edges)
are always done in f32, exactly as the reference does, in both modes; the f32 ln / exp
inside them are the platform libm's logf / expf (what Rust's f32::ln / f32::exp call),
reached through ctypes -- see `_ln`.

Semantics policy: literal everywhere, except where
the literal code is degenerate: `stack_frames` gathers frames[i, j] = x[i*S + j]
(the literal behaviour -- all-zero frames for T > 2 -- is available with literal=True),
`mel_spectrogram1` is `mel_spectrogram2` on one row, `lmfe` is public, `ssc` is defined
here (no reference code exists).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.fft

F32 = np.float32
EPS = np.float32(np.finfo(np.float32).eps)  # f32::EPSILON, functions.rs:70
DCT2_SCALE = 2.0  # ndrustfft default normalisation of nddct2 (feature.rs:123)
CMVN_EPS = 2.0 ** -30  # processing.rs:266


def _rust_round(x: np.float32) -> int:
    """f32::round — half away from zero (processing.rs:77-78)."""
    x = float(x)
    return int(math.floor(abs(x) + 0.5)) * (1 if x >= 0 else -1)


# --------------------------------------------------------------------------- config


@dataclass
class SpeechConfig:
    """config.rs:99-185 — `SpeechConfig::new(sample_rate, fft_points, frame_length,
    frame_stride, num_cepstral, num_filters, low_frequency, high_frequency,
    dc_elimination)`; derived fields as computed at config.rs:151-179."""

    sample_rate: int = 16000
    fft_points: int = 512
    frame_length: float = 0.02
    frame_stride: float = 0.01
    num_cepstral: int = 13
    num_filters: int = 40
    low_frequency: float = 0.0
    high_frequency: float | None = None
    dc_elimination: bool = True
    analysis_mem: np.ndarray = field(default=None, repr=False)  # config.rs:126,162

    def __post_init__(self):
        if self.high_frequency is None:  # SpeechConfigBuilder::new, config.rs:44
            self.high_frequency = float(F32(self.sample_rate) / F32(2.0))
        self.window_size = int(self.fft_points)
        self.window_size_half = self.window_size // 2
        # config.rs:154: (frame_length * sample_rate as f32) as usize  (truncation)
        self.frame_size = int(F32(self.frame_length) * F32(self.sample_rate))
        self.freq_size = self.window_size // 2 + 1  # config.rs:179
        # config.rs:178: 1.0 / (fft_points.pow(2) as f32 / (2 * frame_size) as f32)
        if self.frame_size > 0:
            self.wnorm = F32(1.0) / (F32(self.window_size ** 2) / F32(2 * self.frame_size))
        else:
            self.wnorm = F32(0.0)
        # config.rs:151-160: Vorbis window evaluated in f64, stored as f32
        i = np.arange(self.window_size, dtype=np.float64)
        s = np.sin(0.5 * np.pi * (i + 0.5) / self.window_size_half) if self.window_size_half else i
        self.window = np.sin(0.5 * np.pi * s * s).astype(F32)
        if self.analysis_mem is None and self.window_size >= self.frame_size:
            self.analysis_mem = np.zeros(self.window_size - self.frame_size, F32)  # config.rs:162

    # stack_frames' own derivation (processing.rs:77-78): round, not truncate
    @property
    def frame_sample_length(self) -> int:
        return _rust_round(F32(self.sample_rate) * F32(self.frame_length))

    @property
    def frame_step(self) -> int:
        return _rust_round(F32(self.sample_rate) * F32(self.frame_stride))


# ------------------------------------------------------------------ functions.rs helpers


def _libm_f32(name):
    """logf / expf of the platform libm through ctypes (None if there is no loadable libm)."""
    import ctypes
    import ctypes.util
    try:
        fn = getattr(ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6"), name)
    except (OSError, AttributeError):
        return None
    fn.restype, fn.argtypes = ctypes.c_float, [ctypes.c_float]
    return fn


_LOGF, _EXPF = _libm_f32("logf"), _libm_f32("expf")


def _f32_fn(x, libm_fn, np_fn):
    x = np.asarray(x)
    if x.dtype != F32:
        return np_fn(x)
    if libm_fn is None:  # correctly rounded: equal to glibc's on all but ~0.1 % of the arguments
        return np_fn(x.astype(np.float64)).astype(F32)
    return np.array([libm_fn(float(v)) for v in x.ravel()], F32).reshape(x.shape)


def _ln(x):
    """f32::ln as the reference sees it: Rust lowers f32::ln / f32::exp to the platform libm's logf / expf, so
    the oracle calls those (the band-edge arrays are tiny).  numpy's SIMD float32 loops are 1 ulp off on a few
    percent of the arguments, and that matters: high_freq = sr / 2 puts (F + 1) * hz / sr of the last band
    edge exactly on an integer, where the truncation at feature.rs:70 turns one ulp into a different FFT bin
    (3.7 % of random parameter sets; tests/test_host_abi.py sweeps it)."""
    return _f32_fn(x, _LOGF, np.log)


def _exp(x):
    return _f32_fn(x, _EXPF, np.exp)


def frequency_to_mel(f, dtype=F32):
    """functions.rs:19-21: 1127 * ln(1 + f/700)."""
    f = dtype(f)
    return dtype(1127.0) * _ln(dtype(1.0) + f / dtype(700.0))[()]


def mel_to_frequency(m):
    """functions.rs:36-41: 700 * (exp(m/1127) - 1), elementwise, dtype of m."""
    t = m.dtype.type
    return t(700.0) * (_exp(m / t(1127.0)) - t(1.0))


def triangle(arr, left, middle, right):
    """functions.rs:43-60.  `left <= x < right`; the `middle <= x` branch overrides."""
    out = np.zeros_like(arr)
    for n, x in enumerate(arr):
        if left <= x < right:
            v = out.dtype.type(0.0)
            if x <= middle:
                with np.errstate(divide="ignore", invalid="ignore"):
                    v = (x - left) / (middle - left)
            if middle <= x:
                v = (right - x) / (right - middle)
            out[n] = v
    return out


def zero_handling(x):
    """functions.rs:66-71: exact zeros -> f32::EPSILON."""
    x = np.array(x, copy=True)
    x[x == 0] = x.dtype.type(EPS)
    return x


def filterbank_indices(num_filter, coefficients, sample_rate, low_freq, high_freq):
    """feature.rs:43-70 — always f32: mel linspace, back to Hz, trunc((F+1)*hz/sr)."""
    sr = F32(sample_rate)
    high = F32(sr / F32(2.0)) if high_freq is None else F32(high_freq)
    low = F32(300.0) if low_freq is None else F32(low_freq)  # feature.rs:46
    if not high <= sr / F32(2.0):
        raise ValueError("High frequency cannot be greater than half of the sampling frequency!")
    if not low >= 0:
        raise ValueError("low frequency cannot be less than zero!")
    m0, m1 = frequency_to_mel(low), frequency_to_mel(high)
    n = num_filter + 2
    step = (m1 - m0) / F32(n - 1) if n > 1 else F32(0)  # ndarray linspace: start + step*i
    mels = (m0 + step * np.arange(n, dtype=F32)).astype(F32)
    hz = mel_to_frequency(mels)
    return (F32(coefficients + 1) * hz / sr).astype(np.int64)  # `as usize` truncation, x >= 0


def filterbanks(num_filter, coefficients, sample_rate, low_freq, high_freq, dtype=F32):
    """feature.rs:36-90."""
    idx = filterbank_indices(num_filter, coefficients, sample_rate, low_freq, high_freq)
    fb = np.zeros((num_filter, coefficients), dtype)
    for i in range(num_filter):
        left, middle, right = int(idx[i]), int(idx[i + 1]), int(idx[i + 2])
        z = dtype(left) + np.arange(right - left + 1).astype(dtype)  # linspace, step == 1
        fb[i, left:right + 1] = triangle(z, dtype(left), dtype(middle), dtype(right))
    return fb


# ------------------------------------------------------------------------ processing.rs


def preemphasis(signal, shift=1, cof=0.98, dtype=F32):
    """processing.rs:31-53: signal - cof * np.roll(signal, shift) (circular)."""
    x = np.asarray(signal, dtype)
    n = x.shape[0]
    if shift == 0 or abs(shift) > n:
        raise ValueError("preemphasis: shift must satisfy 0 < |shift| <= len (reference panics)")
    return (x - dtype(cof) * np.roll(x, shift)).astype(dtype)


def num_frames(length_signal, frame_sample_length, frame_step, zero_padding):
    """processing.rs:91-92,101-102: f32 division, ceil (padding) or floor; no '+1'."""
    if length_signal < frame_sample_length:
        raise ValueError("stack_frames: signal shorter than one frame (usize underflow in the reference)")
    q = F32(length_signal - frame_sample_length) / F32(frame_step)
    return int(math.ceil(q)) if zero_padding else int(math.floor(q))


def stack_frames(signal, sample_rate, frame_length, frame_stride, zero_padding=False,
                 literal=False, dtype=F32):
    """processing.rs:65-129 (filter=None).  Default: the intended speechpy gather
    frames[i, j] = x[i*S + j].  literal=True reproduces lines 110-120 as written."""
    x = np.asarray(signal, dtype)
    n = x.shape[0]
    L = _rust_round(F32(sample_rate) * F32(frame_length))
    S = _rust_round(F32(sample_rate) * F32(frame_stride))
    T = num_frames(n, L, S, zero_padding)
    if zero_padding:
        len_sig = T * S + L  # processing.rs:95
        x = np.concatenate([x, np.zeros(len_sig - n, dtype)])
    else:
        if T < 1:
            raise ValueError("stack_frames: fewer than one frame (usize underflow in the reference)")
        len_sig = (T - 1) * S + L  # processing.rs:105
        x = x[:len_sig]
    frames = np.zeros((T, L), dtype)
    if literal:
        # exact_chunks((T, 2)) over the (2, len_sig) doubled signal: 2 // T chunks along
        # axis 0 and len_sig // 2 along axis 1; frames has T // T = 1 by L // 2 chunks.
        doubled = np.stack([x, x])
        n0 = 2 // T if T > 0 else 0
        if n0 > 0:
            for c in range(min(L // 2, len_sig // 2)):
                frames[0:T, 2 * c:2 * c + 2] = doubled[0:T, 2 * c:2 * c + 2]
        return frames
    idx = (np.arange(T)[:, None] * S + np.arange(L)[None, :])
    return x[idx]


def fft_spectrum(frames, fft_points, dtype=F32):
    """processing.rs:143-169: right zero-pad to fft_points, R2C FFT, sqrt(re^2+im^2)."""
    frames = np.asarray(frames, dtype)
    L = frames.shape[1]
    if L > fft_points:
        raise ValueError("fft_spectrum: frame longer than fft_points (ndrustfft panics)")
    if L < fft_points:
        frames = np.pad(frames, [(0, 0), (0, fft_points - L)])
    spec = scipy.fft.rfft(frames, axis=1)  # single precision for f32 input
    return np.sqrt(spec.real ** 2 + spec.imag ** 2).astype(dtype)


def power_spectrum(frames, fft_points, dtype=F32):
    """processing.rs:179-181: (1/fft_points) * |X|  (magnitude, not squared)."""
    return (dtype(1.0) / dtype(fft_points)) * fft_spectrum(frames, fft_points, dtype)


def log_power_spectrum(frames, fft_points, normalize=True, dtype=F32):
    """processing.rs:194-214 (private, unused by the features): 10 log10(power_spectrum) where the value
    exceeds 1e-20, else -200.  `mx` is overwritten by every element above the threshold as `map_mut`
    walks the array in row-major order, so `normalize` subtracts the LAST such value, not the maximum
    (quirk Q7); with no element above the threshold mx stays 1e-20."""
    p = power_spectrum(frames, fft_points, dtype)
    above = p > dtype(1e-20)
    out = np.full(p.shape, dtype(-200.0), dtype)
    out[above] = dtype(10.0) * np.log10(p[above])
    if normalize:
        idx = np.flatnonzero(above.ravel())
        mx = out.ravel()[idx[-1]] if idx.size else dtype(1e-20)
        out = (out - mx).astype(dtype)
    return out


# ---- librosa-compatible mel spectrogram (no reference code: the reference's
# README.md:3,44-46 and Contributing.md:7-8 only name librosa compatibility as the goal).  Restated
# from librosa's documented algorithm (librosa.stft / librosa.filters.mel / librosa.power_to_db) and
# pinned against torchaudio.transforms.MelSpectrogram in tests/test_oracle_kat.py.
def _slaney_hz_to_mel(f):
    f = np.asarray(f, np.float64)
    f_sp, min_log_hz = 200.0 / 3.0, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, f / f_sp)


def _slaney_mel_to_hz(m):
    m = np.asarray(m, np.float64)
    f_sp, min_log_hz = 200.0 / 3.0, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def librosa_mel_filters(sr, n_fft, n_mels, fmin=0.0, fmax=None):
    """librosa.filters.mel(sr=, n_fft=, n_mels=, fmin=, fmax=, htk=False, norm="slaney") -> [n_mels, 1 + n_fft/2] f32"""
    fmax = sr / 2.0 if fmax is None else fmax
    fftfreqs = np.arange(1 + n_fft // 2, dtype=np.float64) * sr / n_fft
    mel_f = _slaney_mel_to_hz(np.linspace(_slaney_hz_to_mel(fmin), _slaney_hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = np.maximum(0.0, np.minimum(lower, upper))
    w *= (2.0 / (mel_f[2:] - mel_f[:-2]))[:, None]
    return w.astype(np.float32)


def librosa_melspectrogram(y, sr, n_fft=2048, hop_length=512, n_mels=128, fmin=0.0, fmax=None, db=False, dtype=F32):
    """librosa.feature.melspectrogram(y=, sr=, n_fft=, hop_length=, window="hann", center=True,
    pad_mode="reflect", power=2.0, n_mels=, fmin=, fmax=) on rows of `y`; `db` applies
    librosa.power_to_db(S, ref=1.0, amin=1e-10, top_db=None).  Returns [..., n_mels, 1 + n / hop]."""
    y = np.asarray(y, dtype)
    one = y.ndim == 1
    y2 = y[None] if one else y
    n = y2.shape[1]
    win = (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n_fft) / n_fft)).astype(dtype)   # periodic Hann
    yp = np.pad(y2, ((0, 0), (n_fft // 2, n_fft // 2)), mode="reflect")
    T = 1 + n // hop_length
    idx = np.arange(T)[:, None] * hop_length + np.arange(n_fft)[None, :]
    frames = yp[:, idx] * win                                       # [rows, T, n_fft]
    X = np.fft.rfft(frames.astype(np.float64), axis=-1)
    S = (X.real ** 2 + X.imag ** 2).astype(dtype)                    # power = 2
    mel = np.einsum("mf,rtf->rmt", librosa_mel_filters(sr, n_fft, n_mels, fmin, fmax).astype(dtype), S).astype(dtype)
    if db:
        mel = (dtype(10.0) * np.log10(np.maximum(mel, dtype(1e-10)))).astype(dtype)
    return mel[0] if one else mel


def cmvn(vec, variance_normalization=False, dtype=F32):
    """processing.rs:265-300.  mean_axis(0): sequential row accumulation then /rows;
    std_axis(0, ddof=0): Welford (ndarray 0.15 `var_axis`); eps = 2^-30."""
    x = np.asarray(vec, dtype)
    rows = x.shape[0]
    if rows == 0:
        raise ValueError("cmvn: empty input (mean_axis(..).unwrap() panics)")
    mean = np.cumsum(x, axis=0, dtype=dtype)[-1] / dtype(rows)
    y = x - mean[None, :]
    if not variance_normalization:
        return y.astype(dtype)
    m = np.zeros(x.shape[1], dtype)
    ssq = np.zeros(x.shape[1], dtype)
    if rows <= 4096:
        for i in range(rows):
            delta = y[i] - m
            m = m + delta / dtype(i + 1)
            ssq = ssq + delta * (y[i] - m)
        var = ssq / dtype(rows)
    else:  # long streams: Welford in a Python loop is too slow; two-pass in the same dtype
        mm = np.cumsum(y, axis=0, dtype=dtype)[-1] / dtype(rows)
        var = np.cumsum((y - mm) ** 2, axis=0, dtype=dtype)[-1] / dtype(rows)
    std = np.sqrt(var)
    return (y / (std[None, :] + dtype(CMVN_EPS))).astype(dtype)


# --------------------------------------------------------------------------- feature.rs


def mfe(signal, cfg: SpeechConfig, dtype=F32, literal_framing=False):
    """feature.rs:200-233 -> (features [T, nf], frame_energies [T])."""
    frames = stack_frames(signal, cfg.sample_rate, cfg.frame_length, cfg.frame_stride,
                          False, literal_framing, dtype)
    P = power_spectrum(frames, cfg.window_size, dtype)
    energy = zero_handling(P.sum(axis=1, dtype=dtype))
    fb = filterbanks(cfg.num_filters, cfg.freq_size, cfg.sample_rate,
                     cfg.low_frequency, cfg.high_frequency, dtype)
    feat = zero_handling((P @ fb.T).astype(dtype))
    return feat, energy


def lmfe(signal, cfg, dtype=F32, literal_framing=False):
    """feature.rs:242-245."""
    return np.log(mfe(signal, cfg, dtype, literal_framing)[0])


def ssc(signal, cfg, dtype=F32):
    """Spectral subband centroids.  No reference code (README.md:15 only); defined in
    SURVEY.md A.9: R = linspace(1, sr/2, F); ssc = ((P*R) @ fb.T) / mfe_features."""
    frames = stack_frames(signal, cfg.sample_rate, cfg.frame_length, cfg.frame_stride, False, False, dtype)
    P = power_spectrum(frames, cfg.window_size, dtype)
    fb = filterbanks(cfg.num_filters, cfg.freq_size, cfg.sample_rate,
                     cfg.low_frequency, cfg.high_frequency, dtype)
    feat = zero_handling((P @ fb.T).astype(dtype))
    F = cfg.freq_size
    step = (dtype(cfg.sample_rate) / dtype(2.0) - dtype(1.0)) / dtype(F - 1)
    R = (dtype(1.0) + step * np.arange(F).astype(dtype)).astype(dtype)
    return (((P * R[None, :]) @ fb.T) / feat).astype(dtype)


def dct2(x, dtype=F32):
    """ndrustfft nddct2 along axis 1 (feature.rs:120-123): DCT2_SCALE * sum x cos(...)."""
    x = np.asarray(x, dtype)
    y = scipy.fft.dct(x, type=2, axis=1, norm=None)  # scipy: 2 * sum x[n] cos(pi k (2n+1)/2N)
    return (y * dtype(DCT2_SCALE / 2.0)).astype(dtype)


def mfcc(signal, cfg: SpeechConfig, dtype=F32, literal_framing=False):
    """feature.rs:99-148."""
    feat, energy = mfe(signal, cfg, dtype, literal_framing)
    if feat.size == 0:
        return np.zeros((0, cfg.num_cepstral), dtype)
    feat = np.log(feat)
    out = dct2(feat, dtype)
    n = dtype(out.size)  # T * num_filters, feature.rs:126
    out[0, 0] *= dtype(1.0) / np.sqrt(dtype(4.0) * n)
    out[:, 1:] *= dtype(1.0) / np.sqrt(dtype(2.0) * n)
    out = out[:, :cfg.num_cepstral].copy()
    if cfg.dc_elimination:
        out[:, 0] = np.log(energy)
    return out.astype(dtype)


# ------------------------------------------------------------------- functions.rs STFT


def stft_num_frames(ttd, frame_size):
    """functions.rs:97 without n_pad: ceil(ttd as f32 / frame_size as f32)."""
    return int(math.ceil(F32(ttd) / F32(frame_size)))


def stft2(signal, cfg: SpeechConfig, dtype=F32):
    """functions.rs:86-170.  Rows share `cfg.analysis_mem` sequentially and the memory
    persists across calls (it is never reset).  Output [ch, T', F] complex; the first
    n_pad frames are dropped, so the last n_pad output rows are exactly zero."""
    x = np.asarray(signal, dtype)
    ch, ttd = x.shape
    W, H, F = cfg.window_size, cfg.frame_size, cfg.freq_size
    if H <= 0 or 2 * H > W:
        raise ValueError("stft: requires 0 < frame_size <= fft_points/2 (usize underflow in the reference)")
    n_pad = W // H - 1
    Tp = stft_num_frames(ttd, H)
    tfd = Tp + n_pad
    cdt = np.complex64 if dtype == F32 else np.complex128
    out = np.zeros((ch, tfd, F), cdt)
    win = cfg.window.astype(dtype)
    mem = cfg.analysis_mem.astype(dtype)
    wnorm = dtype(cfg.wnorm)
    for c in range(ch):
        stream = np.concatenate([mem, x[c], np.zeros(Tp * H - ttd, dtype)])
        # frame j covers stream[(j)*H : j*H + W] of the memory-prefixed stream
        idx = np.arange(Tp)[:, None] * H + np.arange(W)[None, :]
        buf = stream[idx] * win[None, :]
        out[c, :Tp, :] = scipy.fft.rfft(buf, axis=1) * wnorm
        mem = stream[-(W - H):].copy() if W > H else mem
    cfg.analysis_mem = mem.astype(F32)
    return out[:, n_pad:, :]


def stft1(signal, cfg, dtype=F32):
    """functions.rs:199-233."""
    return stft2(np.asarray(signal)[None, :], cfg, dtype)[0]


def mel_spectrogram2(signal, cfg: SpeechConfig, dtype=F32):
    """feature.rs:163-174: |X|^2 via hypot then powi(2); einsum 'ntf,mf->nmt'."""
    X = stft2(signal, cfg, dtype)
    Pw = (np.hypot(X.real, X.imag).astype(dtype)) ** 2
    fb = filterbanks(cfg.num_filters, cfg.freq_size, cfg.sample_rate,
                     cfg.low_frequency, cfg.high_frequency, dtype)
    return np.einsum("ntf,mf->nmt", Pw, fb, optimize=True).astype(dtype)


def mel_spectrogram1(signal, cfg, dtype=F32):
    """feature.rs:151-162, intended semantics (policy D2): one-row mel_spectrogram2."""
    return mel_spectrogram2(np.asarray(signal)[None, :], cfg, dtype)[0]


def log_mel_spectrogram2(signal, cfg, dtype=F32):
    """SURVEY.md A.10 (no reference function): ln(zero_handling(mel_spectrogram2))."""
    return np.log(zero_handling(mel_spectrogram2(signal, cfg, dtype)))


# ---------------------------------------------------------------- synthetic test signals


def synth_signal(n_clips, clip_len, seed=0x5EED0000):
    """N(0, 0.1^2) f32 clips, the reference fixture's distribution (lib.rs:18-22)."""
    rng = np.random.Generator(np.random.Philox(seed))
    return (rng.standard_normal((n_clips, clip_len), dtype=np.float32) * np.float32(0.1))


# ------------------------------------------------- "next" rows of SURVEY 8(f): deltas, cmvnw
#
# The reference's util::pad symmetric / edge modes are degenerate as written (util.rs:75-369:
# `pad` copies the input into the top-left corner and `edge_pad` then overwrites the centre with
# the last row; `symmetric_pad` / `set_edge` assign mismatched shapes), so these two functions
# follow the documented speechpy behaviour the Rust code transliterates: numpy's 'edge' and
# 'symmetric' padding.  The arithmetic (including the `Range * A - B` precedence shared by
# speechpy and processing.rs:244-247) is literal.


def derivative_extraction(feat, delta_windows=2, dtype=F32):
    """processing.rs:222-254: edge-pad along the FEATURE axis, sum_r (r * F[:, c+r] - F[:, c-r]) / (2 sum r^2)."""
    f = np.asarray(feat, dtype)
    cols = f.shape[1]
    dw = int(delta_windows)
    if dw < 1:
        raise ValueError("delta_windows must be >= 1 (division by zero scale in the reference)")
    F = np.pad(f, ((0, 0), (dw, dw)), mode="edge")
    dif = np.zeros_like(f)
    scale = dtype(0.0)
    for i in range(dw):
        r = i + 1
        dif = dif + (F[:, dw + r:dw + r + cols] * dtype(r) - F[:, dw - r:dw - r + cols])
        scale = scale + dtype(2.0) * dtype(r) ** 2
    return (dif / scale).astype(dtype)


def extract_derivative_feature(feat, dtype=F32):
    """feature.rs:253-269: [T, C] -> [T, C, 3] = (static, delta, delta-delta), DeltaWindows = 2."""
    f = np.asarray(feat, dtype)
    d1 = derivative_extraction(f, 2, dtype)
    d2 = derivative_extraction(d1, 2, dtype)
    return np.stack([f, d1, d2], axis=2)


def cmvnw(vec, win_size=301, variance_normalization=False, dtype=F32):
    """processing.rs:315-371: sliding-window CMVN, symmetric padding along time, eps = 2^-30, ddof 0."""
    x = np.asarray(vec, dtype)
    if win_size % 2 != 1:
        raise ValueError("Windows size must be odd!")
    rows = x.shape[0]
    p = (win_size - 1) // 2
    xp = np.pad(x, ((p, p), (0, 0)), mode="symmetric")
    cs = np.concatenate([np.zeros((1, x.shape[1]), np.float64), np.cumsum(xp.astype(np.float64), axis=0)])
    mean = ((cs[win_size:win_size + rows] - cs[:rows]) / win_size).astype(dtype)
    y = (x - mean).astype(dtype)
    if not variance_normalization:
        return y
    yp = np.pad(y, ((p, p), (0, 0)), mode="symmetric").astype(np.float64)
    c1 = np.concatenate([np.zeros((1, x.shape[1])), np.cumsum(yp, axis=0)])
    c2 = np.concatenate([np.zeros((1, x.shape[1])), np.cumsum(yp * yp, axis=0)])
    m = (c1[win_size:win_size + rows] - c1[:rows]) / win_size
    v = np.maximum((c2[win_size:win_size + rows] - c2[:rows]) / win_size - m * m, 0.0)
    return (y / (np.sqrt(v).astype(dtype) + dtype(CMVN_EPS))).astype(dtype)


def cmvnw_direct(vec, win_size=301, variance_normalization=False, dtype=F32):
    """the same function written exactly like the reference loops (O(T * win)); used to pin the
    prefix-sum version above on small inputs"""
    x = np.asarray(vec, dtype)
    rows = x.shape[0]
    p = (win_size - 1) // 2
    xp = np.pad(x, ((p, p), (0, 0)), mode="symmetric")
    y = np.stack([x[i] - xp[i:i + win_size].mean(axis=0, dtype=dtype) for i in range(rows)]).astype(dtype)
    if not variance_normalization:
        return y
    yp = np.pad(y, ((p, p), (0, 0)), mode="symmetric")
    return np.stack([y[i] / (yp[i:i + win_size].std(axis=0, dtype=dtype) + dtype(CMVN_EPS)) for i in range(rows)]).astype(dtype)
