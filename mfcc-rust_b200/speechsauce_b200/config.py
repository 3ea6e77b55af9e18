"""Mirror of speechsauce::config (config.rs:10-190): SpeechConfig / SpeechConfigBuilder.

The config owns an `ss_plan` (device tables + the STFT overlap memory that the reference keeps in
`SpeechConfig::analysis_mem`, config.rs:126).  Like the reference type it is Send but not Sync:
use one config per thread."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, lib


class SpeechConfig:
    """`SpeechConfig::new(sample_rate, fft_points, frame_length, frame_stride, num_cepstral,
    num_filters, low_frequency, high_frequency, dc_elimination)` — config.rs:140-150."""

    def __init__(self, sample_rate=16000, fft_points=512, frame_length=0.02, frame_stride=0.01,
                 num_cepstral=13, num_filters=40, low_frequency=0.0, high_frequency=None,
                 dc_elimination=True, *, devices=None, flags=0, dct2_scale=0.0):
        self.sample_rate = int(sample_rate)
        self.window_size = int(fft_points)
        self.frame_length = float(frame_length)
        self.frame_stride = float(frame_stride)
        self.num_cepstral = int(num_cepstral)
        self.num_filters = int(num_filters)
        self.low_frequency = float(low_frequency)
        self.high_frequency = float(sample_rate) / 2.0 if high_frequency is None else float(high_frequency)
        self.dc_elimination = bool(dc_elimination)
        self._cfg = _lib.ss_config(self.sample_rate, self.window_size, self.frame_length, self.frame_stride,
                                   self.num_cepstral, self.num_filters, self.low_frequency,
                                   self.high_frequency, int(self.dc_elimination), int(flags), float(dct2_scale))
        self._plan = C.c_void_p()
        self.devices = None if devices is None else [int(d) for d in devices]
        if devices is None:          # every visible device (fails loudly when there is none)
            dev_arr, n = None, -1
        else:                        # explicit list; [] = host-only plan (tables / frame counts only)
            dev_arr, n = (C.c_int * len(devices))(*devices), len(devices)
        check(lib.ss_plan_create(C.byref(self._cfg), dev_arr, n, C.byref(self._plan)))
        info = _lib.ss_plan_info()
        check(lib.ss_plan_get_info(self._plan, C.byref(info)))
        self.info = info
        self.frame_size = int(info.hop)            # config.rs:154
        self.freq_size = int(info.freq_size)       # config.rs:179
        self.window_size_half = self.window_size // 2
        self.wnorm = float(info.wnorm)
        self.frame_sample_length = int(info.frame_len)   # processing.rs:77
        self.frame_step = int(info.frame_step)           # processing.rs:78
        self.n_pad = int(info.n_pad)

    @classmethod
    def builder(cls, sample_rate=16000):
        return SpeechConfigBuilder(sample_rate)

    @property
    def plan(self):
        return self._plan

    @property
    def window(self) -> np.ndarray:
        w = np.empty(self.window_size, np.float32)
        check(lib.ss_debug_window(self._plan, _lib.fptr(w)))
        return w

    @property
    def analysis_mem(self) -> np.ndarray:
        m = np.empty(max(0, self.window_size - self.frame_size), np.float32)
        check(lib.ss_plan_get_analysis_mem(self._plan, _lib.fptr(m), m.size))
        return m

    @analysis_mem.setter
    def analysis_mem(self, value):
        if value is None:
            check(lib.ss_plan_set_analysis_mem(self._plan, None, 0))
        else:
            v = np.ascontiguousarray(value, np.float32)
            check(lib.ss_plan_set_analysis_mem(self._plan, _lib.fptr(v), v.size))

    def set_tuning(self, tile_frames=0, chunk_bytes=0):
        """ss_plan_set_tuning: host-pipeline knobs (frame tiles for long clips, bytes per chunk)."""
        check(lib.ss_plan_set_tuning(self._plan, int(tile_frames), int(chunk_bytes)))
        return self

    def filterbank(self) -> np.ndarray:
        fb = np.empty((self.num_filters, self.freq_size), np.float32)
        check(lib.ss_debug_filterbank(self._plan, _lib.fptr(fb)))
        return fb

    def close(self):
        if getattr(self, "_plan", None) is not None and self._plan.value:
            lib.ss_plan_destroy(self._plan)
            self._plan = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SpeechConfigBuilder:
    """config.rs:34-97 (consuming builder; defaults of `SpeechConfigBuilder::new`)."""

    def __init__(self, sample_rate):
        self._kw = dict(sample_rate=sample_rate, fft_points=512, frame_length=0.02, frame_stride=0.01,
                        num_cepstral=13, num_filters=40, low_frequency=0.0,
                        high_frequency=float(sample_rate) / 2.0, dc_elimination=True)

    def high_freq(self, v): self._kw["high_frequency"] = v; return self
    def low_freq(self, v): self._kw["low_frequency"] = v; return self
    def dc_elimination(self, v): self._kw["dc_elimination"] = v; return self
    def num_cepstral(self, v): self._kw["num_cepstral"] = v; return self
    def num_filters(self, v): self._kw["num_filters"] = v; return self  # missing in the reference (Q15)
    def frame_stride(self, v): self._kw["frame_stride"] = v; return self
    def frame_length(self, v): self._kw["frame_length"] = v; return self
    def fft_points(self, v): self._kw["fft_points"] = v; return self

    def build(self, **extra) -> SpeechConfig:
        return SpeechConfig(**self._kw, **extra)
