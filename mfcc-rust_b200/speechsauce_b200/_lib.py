"""ctypes binding of include/speechsauce_cuda.h (the same symbols the Rust crate's ffi.rs binds).

There is no CPU fallback: if the shared library is missing this module raises at import, and
every compute call fails with SpeechsauceError when no CUDA device is visible."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("SPEECHSAUCE_CUDA_LIB", _PKG.parent / "lib" / "libspeechsauce_cuda.so"))


class SpeechsauceError(RuntimeError):
    """A non-zero ss_status.  The Rust wrapper turns the same status into panic!()."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"[ss_status {code}] {msg}")
        self.code = code


SS_OK, SS_ERR_INVALID_ARG, SS_ERR_CUDA, SS_ERR_UNSUPPORTED, SS_ERR_NO_DEVICE, SS_ERR_ALLOC = 0, -1, -2, -3, -4, -5
SS_FLAG_FRAMING_LITERAL, SS_FLAG_FORCE_GENERIC, SS_FLAG_LIBROSA = 0x1, 0x2, 0x4
SS_CMVN_NONE, SS_CMVN_MEAN, SS_CMVN_MEAN_VAR = 0, 1, 2


class ss_config(C.Structure):
    _fields_ = [
        ("sample_rate", C.c_int64), ("fft_points", C.c_int64), ("frame_length", C.c_float),
        ("frame_stride", C.c_float), ("num_cepstral", C.c_int64), ("num_filters", C.c_int64),
        ("low_frequency", C.c_float), ("high_frequency", C.c_float), ("dc_elimination", C.c_int32),
        ("flags", C.c_uint32), ("dct2_scale", C.c_float),
    ]


class ss_plan_info(C.Structure):
    _fields_ = [
        ("frame_len", C.c_int64), ("frame_step", C.c_int64), ("hop", C.c_int64), ("n_fft", C.c_int64),
        ("freq_size", C.c_int64), ("n_pad", C.c_int64), ("num_filters", C.c_int64),
        ("num_cepstral", C.c_int64), ("wnorm", C.c_float), ("n_devices", C.c_int32),
        ("fast_path", C.c_int32), ("fast_stft", C.c_int32),
    ]


class ss_feature_outputs(C.Structure):
    _fields_ = [
        ("mfe", C.c_void_p), ("energy", C.c_void_p), ("lmfe", C.c_void_p), ("ssc", C.c_void_p),
        ("mfcc", C.c_void_p), ("mfcc_cmvn_mode", C.c_int32), ("reserved", C.c_int32),
    ]


if not LIB_PATH.exists():
    raise ImportError(
        f"{LIB_PATH} not found: build it with `make -C {_PKG.parent}` (or __graft_entry__.build()). "
        "speechsauce_b200 has no CPU fallback.")
lib = C.CDLL(str(LIB_PATH))

_i64, _i32, _f, _vp, _fp = C.c_int64, C.c_int, C.c_float, C.c_void_p, C.c_void_p
# every symbol include/speechsauce_cuda.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "ss_device_count": (C.c_int, []),
    "ss_last_error": (C.c_char_p, []),
    "ss_version": (C.c_char_p, []),
    "ss_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "ss_host_free": (C.c_int, [_vp]),
    "ss_host_register": (C.c_int, [_vp, C.c_size_t]),
    "ss_host_unregister": (C.c_int, [_vp]),
    "ss_host_is_pinned": (C.c_int, [_vp]),
    "ss_copy_threads": (C.c_int, []),
    "ss_debug_copy2d": (C.c_int, [_vp, C.c_size_t, _vp, C.c_size_t, C.c_size_t, C.c_size_t]),
    "ss_release_scratch": (C.c_int, []),
    "ss_plan_create": (C.c_int, [C.POINTER(ss_config), C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]),
    "ss_plan_destroy": (None, [_vp]),
    "ss_plan_get_info": (C.c_int, [_vp, C.POINTER(ss_plan_info)]),
    "ss_plan_set_tuning": (C.c_int, [_vp, C.c_int64, C.c_int64]),
    "ss_plan_get_analysis_mem": (C.c_int, [_vp, _fp, _i64]),
    "ss_plan_set_analysis_mem": (C.c_int, [_vp, _fp, _i64]),
    "ss_num_frames": (_i64, [_i64, _i64, _i64, C.c_int]),
    "ss_stft_num_frames": (_i64, [_i64, _i64]),
    "ss_features": (C.c_int, [_vp, _fp, _i64, _i64, _i64, C.POINTER(ss_feature_outputs)]),
    "ss_mfe": (C.c_int, [_vp, _fp, _i64, _i64, _i64, _fp, _fp]),
    "ss_lmfe": (C.c_int, [_vp, _fp, _i64, _i64, _i64, _fp]),
    "ss_mfcc": (C.c_int, [_vp, _fp, _i64, _i64, _i64, _fp, C.c_int]),
    "ss_ssc": (C.c_int, [_vp, _fp, _i64, _i64, _i64, _fp]),
    "ss_stft": (C.c_int, [_vp, _fp, _i64, _i64, _i64, _fp, _fp]),
    "ss_mel_spectrogram": (C.c_int, [_vp, _fp, _i64, _i64, _i64, _fp, _fp, C.c_int]),
    "ss_preemphasis": (C.c_int, [_fp, _fp, _i64, _i64, _f]),
    "ss_preemphasis_batch": (C.c_int, [_fp, _fp, _i64, _i64, _i64, _i64, _f]),
    "ss_stack_frames": (C.c_int, [_fp, _i64, _i64, _i64, C.c_int, C.c_uint32, _fp, C.POINTER(_i64)]),
    "ss_power_spectrum": (C.c_int, [_fp, _i64, _i64, _i64, _fp]),
    "ss_cmvn": (C.c_int, [_fp, _i64, _i64, C.c_int, _fp]),
    "ss_cmvn_batch": (C.c_int, [_fp, _i64, _i64, _i64, C.c_int, _fp]),
    "ss_features_dev": (C.c_int, [_vp, C.c_int, _vp, _fp, _i64, _i64, _i64, C.POINTER(ss_feature_outputs)]),
    "ss_stft_dev": (C.c_int, [_vp, C.c_int, _vp, _fp, _i64, _i64, _i64, _fp, _fp, _fp]),
    "ss_mel_spectrogram_dev": (C.c_int, [_vp, C.c_int, _vp, _fp, _i64, _i64, _i64, _fp, _fp, _fp, C.c_int]),
    "ss_preemphasis_dev": (C.c_int, [_vp, _fp, _fp, _i64, _i64, _i64, _i64, _f]),
    "ss_stack_frames_dev": (C.c_int, [_vp, _fp, _i64, _i64, _i64, C.c_int, C.c_uint32, _fp, _i64]),
    "ss_power_spectrum_dev": (C.c_int, [_vp, _fp, _i64, _i64, _i64, _fp]),
    "ss_log_power_spectrum": (C.c_int, [_fp, _i64, _i64, _i64, C.c_int, _fp]),
    "ss_librosa_melspectrogram": (C.c_int, [_vp, _fp, _i64, _i64, _i64, _fp, C.c_int]),
    "ss_librosa_melspectrogram_dev": (C.c_int, [_vp, C.c_int, _vp, _fp, _i64, _i64, _i64, _fp, C.c_int]),
    "ss_librosa_num_frames": (_i64, [_i64, _i64]),
    "ss_log_power_spectrum_dev": (C.c_int, [_vp, _fp, _i64, _i64, _i64, C.c_int, _fp]),
    "ss_cmvn_dev": (C.c_int, [_vp, _fp, _i64, _i64, _i64, C.c_int, _fp]),
    "ss_derivative_extraction": (C.c_int, [_fp, _i64, _i64, C.c_int, _fp]),
    "ss_extract_derivative_feature": (C.c_int, [_fp, _i64, _i64, _fp]),
    "ss_cmvnw": (C.c_int, [_fp, _i64, _i64, _i64, C.c_int, _fp]),
    "ss_derivative_extraction_dev": (C.c_int, [_vp, _fp, _i64, _i64, C.c_int, _fp]),
    "ss_extract_derivative_feature_dev": (C.c_int, [_vp, _fp, _i64, _i64, _fp]),
    "ss_cmvnw_dev": (C.c_int, [_vp, _fp, _i64, _i64, _i64, _i64, C.c_int, _fp]),
    "ss_debug_filterbank": (C.c_int, [_vp, _fp]),
    "ss_debug_window": (C.c_int, [_vp, _fp]),
    "ss_kernel_launch_count": (_i64, []),
    "ss_last_kernel_name": (C.c_char_p, []),
}
for _name, (_res, _args) in SYMBOLS.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


def check(code: int) -> None:
    if code != 0:
        raise SpeechsauceError(code, lib.ss_last_error().decode("utf-8", "replace"))


def fptr(a):
    """data pointer of a C-contiguous float32/complex64 numpy array (or None)."""
    return None if a is None else a.ctypes.data_as(C.c_void_p)
