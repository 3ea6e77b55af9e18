"""ctypes wrapper of oracle/oracle_c.c (the timed CPU baseline).  Test/bench infrastructure only."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_SO = _DIR / "liboracle_c.so"


class orc_config(C.Structure):
    _fields_ = [("sample_rate", C.c_int64), ("fft_points", C.c_int64), ("frame_length", C.c_float),
                ("frame_stride", C.c_float), ("num_cepstral", C.c_int64), ("num_filters", C.c_int64),
                ("low_frequency", C.c_float), ("high_frequency", C.c_float), ("dc_elimination", C.c_int32),
                ("dct2_scale", C.c_float)]


def build():
    if not _SO.exists() or _SO.stat().st_mtime < (_DIR / "oracle_c.c").stat().st_mtime:
        subprocess.check_call(["make", "-C", str(_DIR), "-s"])


def _load():
    build()
    lib = C.CDLL(str(_SO))
    lib.orc_num_frames.restype = C.c_int64
    lib.orc_num_frames.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_int]
    vp = C.c_void_p
    lib.orc_features.argtypes = [C.POINTER(orc_config), vp, C.c_int64, C.c_int64, vp, vp, vp, vp, vp, C.c_int, C.c_int]
    lib.orc_mel_spectrogram2.argtypes = [C.POINTER(orc_config), vp, C.c_int64, C.c_int64, vp, vp, C.c_int, C.c_int]
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def make_config(sample_rate=16000, fft_points=512, frame_length=0.02, frame_stride=0.01, num_cepstral=13,
                num_filters=40, low_frequency=0.0, high_frequency=None, dc_elimination=True):
    hi = float(sample_rate) / 2.0 if high_frequency is None else high_frequency
    return orc_config(sample_rate, fft_points, frame_length, frame_stride, num_cepstral, num_filters,
                      low_frequency, hi, int(dc_elimination), 2.0)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def max_threads():
    return lib().orc_max_threads()


def features(x, cfg, *, mfe=False, energy=False, lmfe=False, mfcc=False, ssc=False, cmvn_mode=0, n_threads=0):
    x = np.ascontiguousarray(x, np.float32)
    n, clip_len = x.shape
    L = int(np.floor(float(np.float32(cfg.sample_rate) * np.float32(cfg.frame_length)) + 0.5))
    S = int(np.floor(float(np.float32(cfg.sample_rate) * np.float32(cfg.frame_stride)) + 0.5))
    T = lib().orc_num_frames(clip_len, L, S, 0)
    if T < 1:
        raise ValueError("no frames")
    out = {}
    if mfe: out["mfe"] = np.empty((n, T, cfg.num_filters), np.float32)
    if energy: out["energy"] = np.empty((n, T), np.float32)
    if lmfe: out["lmfe"] = np.empty((n, T, cfg.num_filters), np.float32)
    if mfcc: out["mfcc"] = np.empty((n, T, cfg.num_cepstral), np.float32)
    if ssc: out["ssc"] = np.empty((n, T, cfg.num_filters), np.float32)
    rc = lib().orc_features(C.byref(cfg), _p(x), n, clip_len, _p(out.get("mfe")), _p(out.get("energy")),
                            _p(out.get("lmfe")), _p(out.get("mfcc")), _p(out.get("ssc")), cmvn_mode, n_threads)
    if rc:
        raise ValueError("orc_features failed")
    return out


def mel_spectrogram2(x, cfg, *, mem0=None, log=False, n_threads=0):
    x = np.ascontiguousarray(x, np.float32)
    ch, ttd = x.shape
    H = int(np.float32(cfg.frame_length) * np.float32(cfg.sample_rate))
    Tp = int(np.ceil(np.float32(ttd) / np.float32(H)))
    out = np.empty((ch, cfg.num_filters, Tp), np.float32)
    rc = lib().orc_mel_spectrogram2(C.byref(cfg), _p(x), ch, ttd, _p(out), _p(mem0), int(log), n_threads)
    if rc:
        raise ValueError("orc_mel_spectrogram2 failed")
    return out
