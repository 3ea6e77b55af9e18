"""CPU tests of the C-ABI library: it loads, exports every symbol include/speechsauce_cuda.h
declares, its host logic (derived sizes, frame counts, filterbank, window) matches the oracle, and
compute calls fail loudly without a device (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import speechsauce_b200 as sb
import speechsauce_oracle as o
from speechsauce_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent
CONFIGS = {
    "c1": dict(sample_rate=16000, fft_points=512, frame_length=0.025, frame_stride=0.01, num_cepstral=13,
               num_filters=40, low_frequency=0.0, high_frequency=8000.0),
    "default": dict(sample_rate=16000),
    "c3": dict(sample_rate=22050, fft_points=2048, frame_length=512 / 22050, frame_stride=0.01, num_cepstral=13,
               num_filters=128, low_frequency=0.0, high_frequency=11025.0),
    "c4": dict(sample_rate=16000, fft_points=400, frame_length=0.01, frame_stride=0.01, num_cepstral=13,
               num_filters=80, low_frequency=0.0, high_frequency=8000.0),
    "odd": dict(sample_rate=8000, fft_points=256, frame_length=0.03, frame_stride=0.0125, num_cepstral=12,
                num_filters=26, low_frequency=100.0, high_frequency=3800.0),
}


def test_header_symbols_all_exported_and_bound():
    header = (ROOT / "include" / "speechsauce_cuda.h").read_text()
    declared = set(re.findall(r"\b(ss_[a-z0-9_]+)\s*\(", header))
    declared -= {"ss_plan", "ss_config", "ss_status", "ss_plan_info", "ss_feature_outputs"}
    assert len(declared) >= 30
    for name in sorted(declared):
        assert hasattr(sb.lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert b"sm_100a" in sb.lib.ss_version()


@pytest.mark.parametrize("name", list(CONFIGS))
def test_derived_sizes_and_tables_match_oracle(name):
    kw = CONFIGS[name]
    cfg = sb.SpeechConfig(**kw, devices=[])  # host-only plan
    oc = o.SpeechConfig(**kw)
    assert cfg.frame_sample_length == oc.frame_sample_length
    assert cfg.frame_step == oc.frame_step
    assert cfg.frame_size == oc.frame_size
    assert cfg.freq_size == oc.freq_size
    assert np.float32(cfg.wnorm) == np.float32(oc.wnorm)
    np.testing.assert_array_equal(cfg.window, oc.window)
    fb = o.filterbanks(oc.num_filters, oc.freq_size, oc.sample_rate, oc.low_frequency, oc.high_frequency)
    np.testing.assert_array_equal(cfg.filterbank(), fb)
    stft_ok = 0 < oc.frame_size and 2 * oc.frame_size <= oc.window_size
    assert cfg.n_pad == (oc.window_size // oc.frame_size - 1 if stft_ok else -1)


def test_filterbank_matches_oracle_on_random_parameter_sets():
    """600 seeded random (sample rate, FFT length, filter count, band edges): the library's host-built bank is
    identical to the oracle's, bin for bin.  high_freq = sr / 2 (and other round ratios) puts the last band edge
    exactly on an integer before the truncation at feature.rs:70, so this also pins the last-ulp behaviour of
    ln / exp on both sides (see speechsauce_oracle._ln)."""
    rng = np.random.Generator(np.random.Philox(0x5EED0A00))
    for _ in range(600):
        sr = int(rng.choice([8000, 11025, 16000, 22050, 32000, 44100, 48000]))
        W = int(rng.choice([64, 128, 200, 256, 400, 480, 512, 1000, 1024, 2048, 4096]))
        nf = int(rng.integers(1, 129))
        lo = float(rng.choice([0.0, 20.0, 64.0, 133.33, 300.0]))
        hi = float(rng.choice([sr / 2, sr / 2, 0.9 * sr / 2, 0.8 * sr / 2, sr / 3, sr / 4]))
        kw = dict(sample_rate=sr, fft_points=W, frame_length=0.4 * W / sr, frame_stride=0.2 * W / sr,
                  num_cepstral=min(13, nf), num_filters=nf, low_frequency=lo, high_frequency=hi)
        cfg = sb.SpeechConfig(**kw, devices=[])
        fb = o.filterbanks(nf, W // 2 + 1, sr, lo, hi)
        np.testing.assert_array_equal(cfg.filterbank(), fb, err_msg=str(kw))


def test_slaney_filterbank_matches_oracle_on_random_parameter_sets():
    """SS_FLAG_LIBROSA plans: the host-built Slaney bank against the oracle's (itself pinned to torchaudio in
    test_oracle_kat.py) on 200 random (sr, n_fft, n_mels, fmin, fmax).  Both are computed in f64 and stored
    as f32; the two libms differ in the last f64 bit at a few triangle corners, so the comparison allows 1e-7 of
    the bank's peak (most parameter sets are bit-identical)."""
    rng = np.random.Generator(np.random.Philox(0x5EED0A20))
    for _ in range(200):
        sr = int(rng.choice([8000, 16000, 22050, 44100, 48000]))
        n_fft = int(rng.choice([128, 256, 400, 512, 1000, 1024, 2048]))
        n_mels = int(rng.integers(4, 129))
        fmin = float(rng.choice([0.0, 30.0, 125.5]))
        fmax = [None, 0.45 * sr, 0.3 * sr][int(rng.integers(0, 3))]
        cfg = sb.feature.librosa_config(sr, n_fft, max(1, n_fft // 4), n_mels, fmin, fmax, devices=[])
        got, want = cfg.filterbank(), o.librosa_mel_filters(sr, n_fft, n_mels, fmin, fmax)
        assert got.shape == want.shape == (n_mels, n_fft // 2 + 1)
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-7 * float(want.max()), err_msg=f"{sr} {n_fft} {n_mels} {fmin} {fmax}")


def test_derived_sizes_and_frame_counts_on_random_parameter_sets():
    """1,000 seeded random (sample rate, FFT length, frame length / stride on and next to the .5 ties of
    round(sr * length), processing.rs:77-78): frame length, step, hop (trunc, config.rs:154), freq_size, wnorm and
    the Vorbis window equal the oracle's, an invalid combination is rejected by both, and the f32 frame-count
    arithmetic (processing.rs:91-102, functions.rs:97) agrees on random signal lengths."""
    rng = np.random.Generator(np.random.Philox(0x5EED0A10))
    for _ in range(1000):
        sr = int(rng.choice([8000, 11025, 16000, 22050, 32000, 44100, 48000]))
        W = int(rng.choice([64, 128, 200, 256, 400, 480, 512, 1000, 1024, 2048, 4096]))
        L = int(rng.integers(2, W + 1))
        S = int(rng.integers(1, L + 1))
        fl = np.float32((L + float(rng.choice([0.0, 0.5, 0.25, -0.25, 0.4999, 0.5001]))) / sr)
        fs = np.float32((S + float(rng.choice([0.0, 0.5, 0.25, -0.25, 0.4999, 0.5001]))) / sr)
        nf = int(rng.integers(1, 65))
        kw = dict(sample_rate=sr, fft_points=W, frame_length=float(fl), frame_stride=float(fs), num_cepstral=min(13, nf),
                  num_filters=nf, low_frequency=0.0, high_frequency=sr / 2)
        try:
            oc = o.SpeechConfig(**kw)
        except Exception:
            with pytest.raises(sb.SpeechsauceError):
                sb.SpeechConfig(**kw, devices=[])
            continue
        cfg = sb.SpeechConfig(**kw, devices=[])
        assert (cfg.frame_sample_length, cfg.frame_step, cfg.frame_size, cfg.freq_size) == \
               (oc.frame_sample_length, oc.frame_step, oc.frame_size, oc.freq_size), kw
        assert np.float32(cfg.wnorm) == np.float32(oc.wnorm), kw
        np.testing.assert_array_equal(cfg.window, oc.window, err_msg=str(kw))
        Lr, Sr = oc.frame_sample_length, oc.frame_step
        for n in (int(rng.integers(1, 10**7)), Lr, Lr + Sr * int(rng.integers(0, 1000))):
            for zp in (0, 1):
                try:
                    want = o.num_frames(n, Lr, Sr, bool(zp))
                    if not zp and want < 1:
                        want = -1
                except ValueError:
                    want = -1
                assert sb.lib.ss_num_frames(n, Lr, Sr, zp) == want, (n, Lr, Sr, zp)
            if oc.frame_size > 0:
                assert sb.lib.ss_stft_num_frames(n, oc.frame_size) == o.stft_num_frames(n, oc.frame_size), (n, oc.frame_size)


def test_frame_counts_follow_reference_f32_arithmetic():
    for n, L, S in [(16000, 400, 160), (16000, 320, 160), (10**6, 320, 160), (57_600_000, 400, 160), (401, 400, 160),
                    (123457, 400, 160), (2**24 + 3, 400, 160)]:
        for zp in (0, 1):
            try:
                want = o.num_frames(n, L, S, bool(zp))
                if not zp and want < 1:
                    want = -1
            except ValueError:
                want = -1
            assert sb.lib.ss_num_frames(n, L, S, zp) == want, (n, L, S, zp)
    assert sb.lib.ss_num_frames(399, 400, 160, 0) == -1  # usize underflow in the reference
    for ttd, H in [(110250, 512), (480000, 160), (1, 160), (0, 160), (161, 160)]:
        assert sb.lib.ss_stft_num_frames(ttd, H) == o.stft_num_frames(ttd, H)


def test_invalid_configs_are_errors_not_aborts():
    with pytest.raises(sb.SpeechsauceError, match="High frequency"):
        sb.SpeechConfig(16000, high_frequency=9000.0, devices=[])      # feature.rs:47-50
    with pytest.raises(sb.SpeechsauceError, match="low frequency"):
        sb.SpeechConfig(16000, low_frequency=-1.0, devices=[])         # feature.rs:51
    with pytest.raises(sb.SpeechsauceError):
        sb.SpeechConfig(16000, num_cepstral=41, devices=[])            # slice ..num_cepstral out of range


def test_plan_info_reports_the_specialised_kernels_for_the_sizes_they_serve():
    """ss_plan_info.fast_path / fast_stft (host-only plans): which shapes have a register-FFT kernel."""
    assert sb.SpeechConfig(**CONFIGS["c1"], devices=[]).info.fast_path == 1            # n_fft 512, speechpy framing
    assert sb.SpeechConfig(**CONFIGS["c3"], devices=[]).info.fast_stft == 1            # fft_points 2048
    assert sb.SpeechConfig(**CONFIGS["c4"], devices=[]).info.fast_stft == 1            # fft_points 400
    for n_fft, hop in ((256, 80), (512, 160), (1024, 256), (2048, 512), (400, 160)):
        assert sb.feature.librosa_config(16000, n_fft, hop, 40, devices=[]).info.fast_stft == 1, n_fft
        kw = dict(sample_rate=16000, fft_points=n_fft, frame_length=(hop + 0.5) / 16000, frame_stride=0.01, num_cepstral=13, num_filters=26,
                  low_frequency=0.0, high_frequency=8000.0)
        assert sb.SpeechConfig(**kw, devices=[]).info.fast_stft == 1, n_fft
        assert sb.SpeechConfig(**kw, devices=[], flags=_lib.SS_FLAG_FORCE_GENERIC).info.fast_stft == 0
    for n_fft in (300, 768, 4096):                                                      # the any-length kernel
        assert sb.feature.librosa_config(16000, n_fft, n_fft // 4, 40, devices=[]).info.fast_stft == 0, n_fft


def test_no_cpu_fallback_without_device():
    if sb.lib.ss_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    cfg = sb.SpeechConfig(**CONFIGS["c1"], devices=[])
    x = np.zeros(16000, np.float32)
    for fn in (lambda: sb.feature.mfcc(x, cfg), lambda: sb.feature.mfe(x, cfg),
               lambda: sb.processing.preemphasis(x), lambda: sb.processing.cmvn(np.zeros((4, 3), np.float32)),
               lambda: sb.processing.power_spectrum(np.zeros((2, 400), np.float32)),
               lambda: sb.SpeechConfig(16000)):
        with pytest.raises(sb.SpeechsauceError) as ei:
            fn()
        assert ei.value.code == _lib.SS_ERR_NO_DEVICE


def test_argument_errors_precede_device_checks():
    cfg = sb.SpeechConfig(**CONFIGS["c1"], devices=[])
    with pytest.raises(sb.SpeechsauceError) as ei:
        sb.feature.mfcc(np.zeros(300, np.float32), cfg)      # N < L: usize underflow in the reference
    assert ei.value.code == _lib.SS_ERR_INVALID_ARG
    dflt = sb.SpeechConfig(16000, devices=[])                # default config: H = 320 > W/2 (Q13)
    with pytest.raises(sb.SpeechsauceError) as ei:
        sb.functions.stft2(np.zeros((1, 4000), np.float32), dflt)
    assert ei.value.code == _lib.SS_ERR_INVALID_ARG
    with pytest.raises(TypeError):
        sb.feature.mfcc(np.zeros(16000, np.float64), cfg)    # f32 only (README.md:17)
    v = np.zeros((10, 3), np.float32)
    for fn in (lambda: sb.processing.cmvnw(v, 300, False),               # even window: processing.rs:327
               lambda: sb.processing.derivative_extraction(v, 0),        # zero scale: processing.rs:253
               lambda: sb.processing.power_spectrum(np.zeros((2, 400), np.float32), 256),
               lambda: sb.preemphasis(np.zeros(5, np.float32), 0, 0.98)):
        with pytest.raises(sb.SpeechsauceError) as ei:
            fn()
        assert ei.value.code == _lib.SS_ERR_INVALID_ARG


def test_python_mirror_keeps_the_py_speechsauce_signatures():
    """py-speechsauce/speechsauce/__init__.py:37-47, 85-95: `mfcc` and `mel_spectrogram` take (signal,
    sampling_frequency, frame_length=0.020, frame_stride=0.01, num_cepstral=13, num_filters=40, fft_length=512,
    low_frequency=0, high_frequency=None, dc_elimination=True); `cmvn` and `preemphasis` are re-exported from the
    extension module (py-speechsauce/src/lib.rs:208-224)."""
    import inspect
    want = [("signal", inspect.Parameter.empty), ("sampling_frequency", inspect.Parameter.empty), ("frame_length", 0.020),
            ("frame_stride", 0.01), ("num_cepstral", 13), ("num_filters", 40), ("fft_length", 512), ("low_frequency", 0),
            ("high_frequency", None), ("dc_elimination", True)]
    for fn in (sb.mfcc, sb.mel_spectrogram):
        got = [(p.name, p.default) for p in inspect.signature(fn).parameters.values()]
        assert got == want, fn.__name__
    assert [p for p in inspect.signature(sb.cmvn).parameters][:2] == ["vec", "variance_normalization"]
    assert [p for p in inspect.signature(sb.preemphasis).parameters][:3] == ["signal", "shift", "cof"]
    assert inspect.signature(sb.preemphasis).parameters["cof"].default == 0.98      # processing.rs:31-35
    assert set(sb.__all__) >= {"mfcc", "mel_spectrogram", "preemphasis", "cmvn"}


def test_staging_copy_pool_matches_numpy_on_contiguous_strided_and_overlapping_rows():
    """the copy pool that stages pageable caller memory (csrc/ss_host.h): runs without a device"""
    import ctypes as C
    import threading
    rng = np.random.default_rng(5)
    assert sb.lib.ss_copy_threads() >= 0
    cases = [(1, 37, 37, 37), (1, 5_000_003, 5_000_003, 5_000_003),        # one short / one long row (pieces)
             (300, 64000, 64000, 64000),                                  # dense rows (one flat copy)
             (257, 64000, 64148, 64000), (13, 3_000_000, 3_000_016, 3_000_004),  # strided source / destination
             (9, 400_000, 160_000, 400_000),                              # overlapping source rows: frame tiles with a halo
             (1000, 52, 64, 52), (0, 10, 10, 10)]
    for rows, row_bytes, spitch, dpitch in cases:
        src = rng.integers(0, 256, max(1, (rows - 1) * spitch + row_bytes if rows else 1), dtype=np.uint8)
        dst = np.zeros(max(1, (rows - 1) * dpitch + row_bytes if rows else 1), np.uint8)
        _lib.check(sb.lib.ss_debug_copy2d(C.c_void_p(dst.ctypes.data), dpitch, C.c_void_p(src.ctypes.data), spitch, row_bytes, rows))
        for r in (0, rows // 2, rows - 1) if rows else ():
            np.testing.assert_array_equal(dst[r * dpitch:r * dpitch + row_bytes], src[r * spitch:r * spitch + row_bytes])
        if rows and dpitch > row_bytes:
            assert not dst[row_bytes:dpitch].any()        # the gaps between destination rows stay untouched
    # several callers at once (the device workers of a multi-GPU plan share the pool)
    big = [rng.integers(0, 256, 20_000_000, dtype=np.uint8) for _ in range(4)]
    outs = [np.zeros_like(b) for b in big]
    th = [threading.Thread(target=lambda b=b, o_=o_: _lib.check(sb.lib.ss_debug_copy2d(C.c_void_p(o_.ctypes.data), b.size, C.c_void_p(b.ctypes.data), b.size, b.size, 1)))
          for b, o_ in zip(big, outs)]
    [t.start() for t in th]
    [t.join() for t in th]
    for b, o_ in zip(big, outs):
        np.testing.assert_array_equal(b, o_)
    with pytest.raises(sb.SpeechsauceError):
        _lib.check(sb.lib.ss_debug_copy2d(C.c_void_p(outs[0].ctypes.data), 10, C.c_void_p(big[0].ctypes.data), 100, 100, 2))
