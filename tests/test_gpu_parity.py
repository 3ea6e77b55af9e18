"""GPU parity tests: the CUDA path, called through the C-ABI, against the CPU oracle on the same
seeded inputs.  Tolerances are BASELINE.json's: rel 1e-4 on power / mel spectra, abs 1e-3 on
log-mel and MFCC (see conftest.assert_spec_close / assert_log_close)."""
import numpy as np
import pytest

import speechsauce_b200 as sb
import speechsauce_oracle as o
from conftest import RTOL_SPEC, assert_log_close, assert_spec_close
from speechsauce_b200 import _lib
from test_host_abi import CONFIGS

pytestmark = pytest.mark.gpu
GENERIC = _lib.SS_FLAG_FORCE_GENERIC


def special_clips(n=16000):
    t = np.arange(n, dtype=np.float32)
    noise = o.synth_signal(1, n, 99)[0]
    sine = (0.5 * np.sin(2 * np.pi * 440.0 * t / 16000.0)).astype(np.float32) + np.float32(0.01) * noise
    imp = np.zeros(n, np.float32)
    imp[1234] = 1.0
    return {"noise": noise, "sine": sine, "zero": np.zeros(n, np.float32), "impulse": imp,
            "dc": np.full(n, 0.25, np.float32)}


@pytest.mark.parametrize("flags", [0, GENERIC], ids=["fast", "generic"])
@pytest.mark.parametrize("cname", ["c1", "default", "odd"])
def test_speechpy_features_match_oracle(cname, flags):
    kw = CONFIGS[cname]
    cfg = sb.SpeechConfig(**kw, flags=flags)
    oc = o.SpeechConfig(**kw)
    if flags == 0 and cname in ("c1", "default"):
        assert cfg.info.fast_path == 1
    n = oc.sample_rate
    for name, x in special_clips(n).items():
        r = sb.feature.features(x, cfg, mfe=True, energy=True, lmfe=True, ssc=True, mfcc=True)
        f, e = o.mfe(x, oc)
        f64, e64 = o.mfe(x, oc, np.float64)
        assert r["mfe"].shape == f.shape and r["energy"].shape == e.shape
        assert_spec_close(r["mfe"], f, what=f"{cname}/{name} mfe")
        assert_spec_close(r["energy"], e, what=f"{cname}/{name} energy")
        assert_spec_close(r["mfe"], f64, what=f"{cname}/{name} mfe vs f64")
        assert_log_close(r["lmfe"], np.log(f), what=f"{cname}/{name} lmfe")
        assert_log_close(r["mfcc"], o.mfcc(x, oc), what=f"{cname}/{name} mfcc")
        assert_spec_close(r["ssc"], o.ssc(x, oc), what=f"{cname}/{name} ssc")
        if name == "zero":  # zero_handling, functions.rs:66-71
            assert (r["mfe"] == o.EPS).all() and (r["energy"] == o.EPS).all()


@pytest.mark.parametrize("flags", [0, GENERIC], ids=["fast", "generic"])
def test_individual_entry_points_and_dc_elimination_off(flags):
    kw = dict(CONFIGS["c1"], dc_elimination=False)
    cfg = sb.SpeechConfig(**kw, flags=flags)
    oc = o.SpeechConfig(**kw)
    x = o.synth_signal(3, 16000, 5)
    m = sb.feature.mfcc(x, cfg)
    assert m.shape == (3, 97, 13)
    for i in range(3):
        assert_log_close(m[i], o.mfcc(x[i], oc), atol=2e-3, what="mfcc dc_elimination=False")  # column 0 is unscaled (|c0| ~ 1e3)
        assert_log_close(m[i][:, 1:], o.mfcc(x[i], oc)[:, 1:], what="mfcc dc_elimination=False cols 1..")
    f, e = sb.feature.mfe(x[0], cfg)
    assert_spec_close(f, o.mfe(x[0], oc)[0], what="ss_mfe")
    assert_log_close(sb.feature.lmfe(x[0], cfg), o.lmfe(x[0], oc), what="ss_lmfe")
    assert_spec_close(sb.feature.ssc(x[0], cfg), o.ssc(x[0], oc), what="ss_ssc")


@pytest.mark.parametrize("flags", [0, GENERIC], ids=["fast", "generic"])
def test_kws_batch_slice_lmfe_mfcc_cmvn(flags):
    """config c2 on a 300-clip slice: lmfe + mfcc with per-clip cmvn(mean, var)."""
    cfg = sb.SpeechConfig(**CONFIGS["c1"], flags=flags)
    oc = o.SpeechConfig(**CONFIGS["c1"])
    x = o.synth_signal(300, 16000, 0x5EED0002)
    for mode in (1, 2):
        r = sb.feature.features(x, cfg, lmfe=True, mfcc=True, cmvn_mode=mode)
        for i in (0, 1, 17, 150, 299):
            assert_log_close(r["lmfe"][i], o.lmfe(x[i], oc), what=f"lmfe clip {i}")
            want = o.cmvn(o.mfcc(x[i], oc), mode == 2)
            assert_log_close(r["mfcc"][i], want, what=f"cmvn(mfcc) clip {i} mode {mode}")
        np.testing.assert_allclose(r["mfcc"].mean(1), 0, atol=2e-5 if mode == 2 else 1e-5)
        if mode == 2:
            np.testing.assert_allclose(r["mfcc"].std(1), 1, atol=1e-4)


@pytest.mark.parametrize("flags", [0, GENERIC], ids=["fast", "generic"])
def test_silence_next_to_sound_has_no_crosstalk(flags):
    """An all-zero frame next to a loud one must still give exactly EPS (zero_handling), and a very
    quiet frame next to a loud one must keep its own relative accuracy.  (Packing two real frames
    into one complex FFT fails both; the kernels use one real FFT per frame.)"""
    cfg = sb.SpeechConfig(**CONFIGS["c1"], flags=flags)
    oc = o.SpeechConfig(**CONFIGS["c1"])
    loud = o.synth_signal(1, 16000, 41)[0] * np.float32(10.0)
    gated = loud.copy()
    gated[:8000] = 0.0                      # digital silence, then sound
    quiet = loud.copy()
    quiet[:8000] *= np.float32(1e-6)        # -120 dB background, then sound
    for name, x in (("gated", gated), ("quiet", quiet)):
        r = sb.feature.features(x, cfg, mfe=True, energy=True, lmfe=True, mfcc=True)
        f, e = o.mfe(x, oc)
        silent = ~frames_touching(x, 400, 160, 97)
        if name == "gated":
            assert silent.sum() > 40
            assert (r["mfe"][silent] == o.EPS).all() and (r["energy"][silent] == o.EPS).all()
        assert_spec_close(r["mfe"], f, what=f"{name} mfe")
        assert_spec_close(r["energy"], e, what=f"{name} energy")
        # per-frame relative accuracy: the tolerance floor scales with that frame, not the clip
        for t in range(97):
            assert_spec_close(r["mfe"][t], f[t], what=f"{name} mfe frame {t}")
        assert_log_close(r["lmfe"], np.log(f), what=f"{name} lmfe")
        assert_log_close(r["mfcc"], o.mfcc(x, oc), what=f"{name} mfcc")


def frames_touching(x, L, S, T):
    """True for frames that contain at least one non-zero sample"""
    nz = np.flatnonzero(x)
    out = np.zeros(T, bool)
    for t in range(T):
        out[t] = ((nz >= t * S) & (nz < t * S + L)).any()
    return out


def test_fast_and_generic_paths_agree_on_ragged_shapes():
    for clip_len, L_s, S_s in [(16000, 0.025, 0.01), (5003, 0.02, 0.01), (1237, 0.032, 0.005), (640, 0.025, 0.01), (561, 0.025, 0.01)]:
        kw = dict(CONFIGS["c1"], frame_length=L_s, frame_stride=S_s)
        a = sb.SpeechConfig(**kw)
        b = sb.SpeechConfig(**kw, flags=GENERIC)
        oc = o.SpeechConfig(**kw)
        assert a.info.fast_path == 1 and b.info.fast_path == 0
        x = o.synth_signal(7, clip_len, clip_len)
        ra = sb.feature.features(x, a, mfe=True, energy=True, mfcc=True, lmfe=True)
        rb = sb.feature.features(x, b, mfe=True, energy=True, mfcc=True, lmfe=True)
        T = o.num_frames(clip_len, oc.frame_sample_length, oc.frame_step, False)
        assert ra["mfe"].shape == (7, T, 40)
        assert_spec_close(ra["mfe"], rb["mfe"], what=f"fast vs generic mfe {clip_len}")
        assert_spec_close(ra["energy"], rb["energy"], what="fast vs generic energy")
        assert_log_close(ra["mfcc"], rb["mfcc"], what="fast vs generic mfcc")
        assert_log_close(ra["mfcc"][3], o.mfcc(x[3], oc), what="fast mfcc vs oracle")


@pytest.mark.parametrize("nf,nc,dc", [(26, 13, True), (40, 16, True), (41, 8, False), (64, 13, True), (17, 12, True)])
def test_fast512_filterbank_and_cepstrum_variants(nf, nc, dc):
    """The n_fft = 512 tensor-core epilogue for other bank / cepstrum sizes than the headline's 40 / 13:
    partial 16-filter m-tiles, an odd filter count, a fourth m-tile (nf > 48), 16 cepstra, c0 scaling."""
    kw = dict(CONFIGS["c1"], num_filters=nf, num_cepstral=nc, dc_elimination=dc, frame_length=0.02)
    cfg = sb.SpeechConfig(**kw)
    oc = o.SpeechConfig(**kw)
    assert cfg.info.fast_path == 1
    clips = special_clips(16000)
    x = np.stack([clips["noise"], clips["sine"], clips["impulse"]])
    r = sb.feature.features(x, cfg, mfe=True, energy=True, lmfe=True, ssc=True, mfcc=True)
    assert sb.lib.ss_last_kernel_name() == b"k_fast512c"
    for i, name in enumerate(("noise", "sine", "impulse")):
        f, e = o.mfe(x[i], oc)
        assert_spec_close(r["mfe"][i], f, what=f"nf {nf} {name} mfe")
        assert_spec_close(r["energy"][i], e, what=f"nf {nf} {name} energy")
        assert_log_close(r["lmfe"][i], np.log(f), what=f"nf {nf} {name} lmfe")
        assert_spec_close(r["ssc"][i], o.ssc(x[i], oc), what=f"nf {nf} {name} ssc")
        want = o.mfcc(x[i], oc)
        if dc:
            assert_log_close(r["mfcc"][i], want, what=f"nf {nf} nc {nc} {name} mfcc")
        else:  # column 0 is unscaled (|c0| ~ 1e3): relative tolerance there
            assert_log_close(r["mfcc"][i][:, 1:], want[:, 1:], what=f"nf {nf} nc {nc} {name} mfcc cols 1..")
            np.testing.assert_allclose(r["mfcc"][i][:, 0], want[:, 0], rtol=2e-6, atol=1e-3)


def test_fast512_fused_cmvn_on_a_full_grid():
    """With at least one clip per resident warp the n_fft = 512 kernel normalises each clip in place
    (one warp per clip, no separate cmvn launch): 3,200 device-resident clips of config c2, both cmvn
    modes, through the *_dev entry point the throughput bench uses."""
    import ctypes as C

    import torch
    cfg = sb.SpeechConfig(**CONFIGS["c1"], devices=[0])
    oc = o.SpeechConfig(**CONFIGS["c1"])
    x = o.synth_signal(3200, 16000, 0x5EED0102)
    x[5, :6000] = 0.0      # a clip that starts with digital silence
    x[6] *= np.float32(30.0)
    dx = torch.from_numpy(x).cuda()
    lm = torch.empty((3200, 97, 40), dtype=torch.float32, device="cuda")
    mf = torch.empty((3200, 97, 13), dtype=torch.float32, device="cuda")
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for mode in (1, 2):
        outs = _lib.ss_feature_outputs(None, None, C.c_void_p(lm.data_ptr()), None, C.c_void_p(mf.data_ptr()), mode, 0)
        n0 = sb.lib.ss_kernel_launch_count()
        _lib.check(sb.lib.ss_features_dev(cfg.plan, 0, sp, C.c_void_p(dx.data_ptr()), 3200, 16000, 16000, C.byref(outs)))
        torch.cuda.synchronize()
        assert sb.lib.ss_kernel_launch_count() - n0 == 1, "cmvn must be fused into the feature kernel"
        assert sb.lib.ss_last_kernel_name() == b"k_fast512c"
        lmfe, mfcc = lm.cpu().numpy(), mf.cpu().numpy()
        for i in (0, 5, 6, 1599, 3199):
            assert_log_close(lmfe[i], o.lmfe(x[i], oc), what=f"lmfe clip {i}")
            assert_log_close(mfcc[i], o.cmvn(o.mfcc(x[i], oc), mode == 2), what=f"cmvn(mfcc) clip {i} mode {mode}")
        np.testing.assert_allclose(mfcc.mean(1), 0, atol=2e-5 if mode == 2 else 1e-5)
        if mode == 2:
            np.testing.assert_allclose(mfcc.std(1), 1, atol=1e-4)


def test_fast512_odd_strides_tiny_clips_and_padded_rows():
    """Shapes that switch off the vector / cp.async paths of k_fast512c or leave partial tiles: an odd
    frame step and length, clips of 1..9 frames, and rows with a stride that is not the clip length
    (straight through the C ABI)."""
    import ctypes as C
    kw = dict(CONFIGS["c1"], frame_length=0.0203125, frame_stride=0.0100625)   # L = 325, S = 161
    cfg, oc = sb.SpeechConfig(**kw), o.SpeechConfig(**kw)
    assert cfg.info.fast_path == 1 and (oc.frame_sample_length, oc.frame_step) == (325, 161)
    x = o.synth_signal(3, 16001, 0x5EED0201)
    r = sb.feature.features(x, cfg, mfe=True, energy=True, lmfe=True, mfcc=True)
    assert sb.lib.ss_last_kernel_name() == b"k_fast512c"
    for i in range(3):
        f, e = o.mfe(x[i], oc)
        assert_spec_close(r["mfe"][i], f, what="odd stride mfe")
        assert_spec_close(r["energy"][i], e, what="odd stride energy")
        assert_log_close(r["mfcc"][i], o.mfcc(x[i], oc), what="odd stride mfcc")
    cfg, oc = sb.SpeechConfig(**CONFIGS["c1"]), o.SpeechConfig(**CONFIGS["c1"])
    for T in (1, 2, 3, 7, 8, 9):
        n = 400 + 160 * T + 17          # floor((n - 400) / 160) = T frames
        y = o.synth_signal(2, n, 0x5EED0210 + T)
        assert o.num_frames(n, 400, 160, False) == T
        r = sb.feature.features(y, cfg, mfe=True, lmfe=True, mfcc=True, cmvn_mode=2)
        assert r["mfcc"].shape == (2, T, 13)
        assert_spec_close(r["mfe"][1], o.mfe(y[1], oc)[0], what=f"T={T} mfe")
        if T > 1:  # cmvn of a single row is 0 / (0 + eps)
            assert_log_close(r["mfcc"][1], o.cmvn(o.mfcc(y[1], oc), True), atol=2e-3, what=f"T={T} cmvn(mfcc)")
    # rows 16,003 floats apart (neither 8- nor 16-byte aligned), 16,000 samples each
    stride = 16003
    buf = np.zeros((4, stride), np.float32)
    buf[:, :16000] = o.synth_signal(4, 16000, 0x5EED0220)
    buf[:, 16000:] = np.float32(1e9)    # must never be read into a frame
    lm = np.empty((4, 97, 40), np.float32)
    mf = np.empty((4, 97, 13), np.float32)
    outs = _lib.ss_feature_outputs(None, None, _lib.fptr(lm), None, _lib.fptr(mf), 0, 0)
    _lib.check(sb.lib.ss_features(cfg.plan, _lib.fptr(buf), 4, 16000, stride, C.byref(outs)))
    for i in range(4):
        assert_log_close(lm[i], o.lmfe(buf[i, :16000], oc), what="strided rows lmfe")
        assert_log_close(mf[i], o.mfcc(buf[i, :16000], oc), what="strided rows mfcc")


def test_fast512_long_stream_chunks():
    """Config c5's shape in small: two 90 s streams are cut into chunks of whole tiles, one per warp."""
    cfg = sb.SpeechConfig(**CONFIGS["c1"])
    gen = sb.SpeechConfig(**CONFIGS["c1"], flags=GENERIC)
    oc = o.SpeechConfig(**CONFIGS["c1"])
    x = o.synth_signal(2, 16000 * 90 + 77, 0x5EED0105)
    r = sb.feature.features(x, cfg, mfe=True, energy=True, ssc=True, mfcc=True)
    g = sb.feature.features(x, gen, mfe=True, energy=True, ssc=True, mfcc=True)
    T = o.num_frames(x.shape[1], 400, 160, False)
    assert r["mfe"].shape == (2, T, 40) and r["mfcc"].shape == (2, T, 13)
    assert_spec_close(r["mfe"], g["mfe"], what="long stream mfe vs generic")
    assert_spec_close(r["energy"], g["energy"], what="long stream energy vs generic")
    assert_spec_close(r["ssc"], g["ssc"], what="long stream ssc vs generic")
    assert_log_close(r["mfcc"], g["mfcc"], what="long stream mfcc vs generic")
    f, e = o.mfe(x[1], oc)
    assert_spec_close(r["mfe"][1], f, what="long stream mfe vs oracle")
    assert_log_close(r["mfcc"][1], o.mfcc(x[1], oc), what="long stream mfcc vs oracle")


@pytest.mark.parametrize("n,length", [(6000, 400 + 160 * 3 + 5), (4097, 400 + 160 * 11), (700, 400 + 160 * 29 + 150)])
def test_unit_runs_across_many_short_rows(n, length):
    """Every resident warp takes one contiguous run of units; with thousands of rows of a few frames each a run
    crosses many row boundaries (and, with fewer units than warps, some runs are empty).  k_fast512c (tiles
    of 8 frames, no fused CMVN) and k_stft400m (24-frame units, 1..2 per row) against the generic kernel,
    bit-for-bit against the same rows computed alone, and the oracle on sampled rows."""
    x = o.synth_signal(n, length, 0x5EED0500 + n)
    cfg, gen, oc = sb.SpeechConfig(**CONFIGS["c1"]), sb.SpeechConfig(**CONFIGS["c1"], flags=GENERIC), o.SpeechConfig(**CONFIGS["c1"])
    r = sb.feature.features(x, cfg, mfe=True, energy=True, mfcc=True)
    assert sb.lib.ss_last_kernel_name() == b"k_fast512c"
    g = sb.feature.features(x, gen, mfe=True, energy=True, mfcc=True)
    assert_spec_close(r["mfe"], g["mfe"], what="short rows mfe vs generic")
    assert_spec_close(r["energy"], g["energy"], what="short rows energy vs generic")
    assert_log_close(r["mfcc"], g["mfcc"], what="short rows mfcc vs generic")
    pick = [0, 1, n // 3, n - 2, n - 1]
    alone = sb.feature.features(x[pick], cfg, mfe=True, energy=True, mfcc=True)
    for k in ("mfe", "energy", "mfcc"):
        assert np.array_equal(alone[k], r[k][pick]), k
    for i in pick[:3]:
        assert_spec_close(r["mfe"][i], o.mfe(x[i], oc)[0], what=f"short row {i} mfe vs oracle")
        assert_log_close(r["mfcc"][i], o.mfcc(x[i], oc), what=f"short row {i} mfcc vs oracle")
    c4, g4, o4 = sb.SpeechConfig(**CONFIGS["c4"]), sb.SpeechConfig(**CONFIGS["c4"], flags=GENERIC), o.SpeechConfig(**CONFIGS["c4"])
    a = sb.feature.mel_spectrogram2(x, c4, log=False)
    assert sb.lib.ss_last_kernel_name() == b"k_stft400m"
    b = sb.feature.mel_spectrogram2(x, g4, log=False)
    # per-row tolerance floor: a band that is empty but for the f32 FFT's own rounding (1e-10 of the row's
    # peak in a handful of entries here) carries different noise in the two kernels and in the oracle
    peak = b.max(axis=(1, 2), keepdims=True).astype(np.float64)
    err = np.abs(a.astype(np.float64) - b)
    assert (err <= RTOL_SPEC * np.abs(b) + 2e-6 * peak).all(), f"short rows mel vs generic: worst {(err / peak).max():.3e} of the row peak"
    al = sb.feature.mel_spectrogram2(x, c4, log=True)
    bl = sb.feature.mel_spectrogram2(x, g4, log=True)
    loud = b > 1e-6 * peak
    assert loud.mean() > 0.6           # 18 of the 80 filters are empty and the last frame of a row is the zero tail
    assert_log_close(al[loud], bl[loud], what="short rows log-mel vs generic")
    want = o.mel_spectrogram2(x[:3], o4)          # rows 0..2 of a fresh plan: the overlap memory starts at zero
    assert_spec_close(a[:3], want, what="short rows mel vs oracle")
    assert_log_close(al[:3][loud[:3]], np.log(want.astype(np.float64)[loud[:3]]), what="short rows log-mel vs oracle")


def test_cuda_path_matches_the_committed_golden_fixtures():
    """tests/golden/*.npz (generated by tests/golden/make_golden.py, pinned against oracle drift by
    test_oracle_kat.py): the CUDA path reproduces them from the same seeded inputs, one fixture per
    BASELINE config shape (c1, a c2 slice, c3, c4)."""
    import json
    from pathlib import Path
    gold = Path(__file__).parent / "golden"
    meta = json.loads((gold / "golden.json").read_text())
    for name, m in meta.items():
        z = np.load(gold / f"{name}.npz")
        cfg = sb.SpeechConfig(**m["config"])
        x = o.synth_signal(m["n_clips"], m["clip_len"], m["seed"])
        if m["kind"] == "speechpy":
            r = sb.feature.features(x, cfg, mfe=True, energy=True, mfcc=True)
            assert_spec_close(r["mfe"], z["mfe"], what=f"{name} mfe")
            assert_spec_close(r["energy"], z["energy"], what=f"{name} energy")
            assert_log_close(r["mfcc"], z["mfcc"], what=f"{name} mfcc")
        else:
            assert_spec_close(sb.feature.mel_spectrogram2(x, cfg), z["mel"], what=f"{name} mel")


def test_reference_shape_tests_1M_samples():
    """speechsauce/src/lib.rs:94-134 re-expressed: default config, 1e6 samples -> (6248, 13), no NaN."""
    sig = o.synth_signal(1, 1_000_000, 7)[0]
    m = sb.mfcc(sig, 16000)  # the py-speechsauce call (py-speechsauce/speechsauce/__init__.py:37)
    assert m.shape == (6248, 13) and not np.isnan(m).any()
    cfg = sb.SpeechConfigBuilder(16000).build()
    f, e = sb.feature.mfe(sig, cfg)
    assert f.shape == (6248, 40) and e.shape == (6248,)
    oc = o.SpeechConfig(16000)
    assert_log_close(m, o.mfcc(sig, oc), what="1M-sample mfcc")
    assert_spec_close(f, o.mfe(sig, oc)[0], what="1M-sample mfe")


def test_literal_framing_flag():
    cfg = sb.SpeechConfig(**CONFIGS["c1"], flags=_lib.SS_FLAG_FRAMING_LITERAL)
    oc = o.SpeechConfig(**CONFIGS["c1"])
    x = o.synth_signal(1, 16000, 3)[0]
    assert_log_close(sb.feature.mfcc(x, cfg), o.mfcc(x, oc, literal_framing=True), what="literal mfcc")
    short = x[:720]  # T == 2: both rows are x[0..L]
    assert_spec_close(sb.feature.mfe(short, cfg)[0], o.mfe(short, oc, literal_framing=True)[0], what="literal T=2")


def test_feature_errors():
    cfg = sb.SpeechConfig(**CONFIGS["c1"])
    with pytest.raises(sb.SpeechsauceError) as ei:
        sb.feature.mfcc(np.zeros(300, np.float32), cfg)
    assert ei.value.code == _lib.SS_ERR_INVALID_ARG
    with pytest.raises(sb.SpeechsauceError):
        sb.feature.mfcc(np.zeros(500, np.float32), cfg)  # floor((500-400)/160) == 0 frames


# ------------------------------------------------------------------ standalone processing functions


def test_preemphasis():
    got = sb.preemphasis(np.array([1, 2, 3, 4, 5], np.float32), 1, 0.98)
    np.testing.assert_allclose(got, [1 - 4.9, 2 - 0.98, 3 - 1.96, 4 - 2.94, 5 - 3.92], rtol=1e-6)
    x = o.synth_signal(1, 1_000_003, 1)[0]
    for shift in (1, 2, -3, 1_000_003):
        np.testing.assert_array_equal(sb.preemphasis(x, shift, 0.98), o.preemphasis(x, shift, 0.98))
    xb = o.synth_signal(5, 4001, 2)
    np.testing.assert_array_equal(sb.preemphasis(xb, 1, 0.97), np.stack([o.preemphasis(r, 1, 0.97) for r in xb]))
    with pytest.raises(sb.SpeechsauceError):
        sb.preemphasis(x, 0, 0.98)


def test_stack_frames_vector_path():
    """L, S multiples of 4: the 16-byte kernel, including the zero-padded tail and a long signal"""
    rng = np.random.default_rng(21)
    for n, L, S, zp in [(16000, 400, 160, False), (16000, 400, 160, True), (1_000_004, 512, 256, True), (1_000_003, 512, 256, True), (5000, 64, 4, False), (776, 400, 160, True)]:
        x = rng.standard_normal(n).astype(np.float32)
        got = sb.processing.stack_frames(x, 16000, L / 16000, S / 16000, None, zp)
        want = o.stack_frames(x, 16000, L / 16000, S / 16000, zp)
        np.testing.assert_array_equal(got, want)


def test_stack_frames():
    x = o.synth_signal(1, 100_000, 4)[0]
    for fl, fs, zp in [(0.02, 0.02, True), (0.025, 0.01, False), (0.025, 0.01, True), (0.02, 0.01, False)]:
        np.testing.assert_array_equal(sb.processing.stack_frames(x, 16000, fl, fs, None, zp),
                                      o.stack_frames(x, 16000, fl, fs, zp))
    np.testing.assert_array_equal(sb.processing.stack_frames(x, 16000, 0.025, 0.01, None, False, literal=True),
                                  o.stack_frames(x, 16000, 0.025, 0.01, False, literal=True))
    np.testing.assert_array_equal(sb.processing.stack_frames(x[:720], 16000, 0.025, 0.01, None, False, literal=True),
                                  o.stack_frames(x[:720], 16000, 0.025, 0.01, False, literal=True))
    with pytest.raises(sb.SpeechsauceError):
        sb.processing.stack_frames(x[:100], 16000, 0.025, 0.01, None, False)


def test_power_spectrum():
    P = sb.processing.power_spectrum(np.ones((1, 320), np.float32), 512)
    k = np.arange(1, 257)
    want = np.abs(np.sin(np.pi * 320 * k / 512) / np.sin(np.pi * k / 512)) / 512
    assert abs(P[0, 0] - 0.625) < 1e-6
    np.testing.assert_allclose(P[0, 1:], want, rtol=1e-4, atol=2e-6)
    fr = o.synth_signal(33, 400, 8)
    for nfft in (512, 400, 1024, 2048, 480):
        assert_spec_close(sb.processing.power_spectrum(fr, nfft), o.power_spectrum(fr, nfft), what=f"power_spectrum {nfft}")
    # n_fft = 512 runs on the register-FFT kernel (k_pspec512): every pruning case of the frame length, odd row counts
    for L in (1, 2, 31, 32, 33, 257, 320, 399, 400, 401, 511, 512):
        f = o.synth_signal(7, L, 100 + L)
        f[3] = 0.0
        assert_spec_close(sb.processing.power_spectrum(f, 512), o.power_spectrum(f, 512), what=f"power_spectrum L={L}")
    assert sb.lib.ss_last_kernel_name() == b"k_pspec512"
    with pytest.raises(sb.SpeechsauceError):
        sb.processing.power_spectrum(fr, 256)  # frame longer than the FFT: ndrustfft panics


def test_log_power_spectrum():
    """processing::log_power_spectrum (processing.rs:194-214), including the last-value normaliser (Q7)."""
    fr = o.synth_signal(33, 400, 8)
    fr[7] = 0.0                                   # a silent frame: -200 everywhere
    fr[32, 1:] = 0.0                              # the last row is an impulse: flat spectrum
    for nfft, norm in [(512, False), (512, True), (400, True), (1024, True)]:
        got = sb.processing.log_power_spectrum(fr, nfft, norm)
        want = o.log_power_spectrum(fr, nfft, norm)
        assert got.shape == want.shape
        assert (got[7] == want[7]).all()
        # 10 log10 turns the rel 1e-4 spectrum tolerance into 4.3e-4 dB; bins that are zero only by
        # cancellation carry the FFT's own noise in both implementations
        loud = o.power_spectrum(fr, nfft) > 1e-6
        assert_log_close(got[loud], want[loud], what=f"log_power_spectrum {nfft} normalize={norm}")
    silent = sb.processing.log_power_spectrum(np.zeros((4, 64), np.float32), 64, True)
    assert (silent == -200.0).all()


def test_librosa_mode_matches_oracle():
    """SS_FLAG_LIBROSA: Hann window, centred reflect-padded frames, Slaney mel bank, |X|^2, optional
    power_to_db; checked against the oracle that tests/test_oracle_kat.py pins to torchaudio."""
    for sr, n_fft, hop, n_mels, fmin, fmax, n in [(22050, 2048, 512, 128, 0.0, None, 22050 + 37),
                                                  (16000, 400, 160, 80, 0.0, 8000.0, 16000),
                                                  (16000, 512, 128, 40, 50.0, 7000.0, 3001), (8000, 256, 256, 26, 0.0, None, 129)]:
        cfg = sb.feature.librosa_config(sr, n_fft, hop, n_mels, fmin, fmax)
        assert cfg.frame_size == hop
        y = o.synth_signal(3, n, 0x5EED0310 + n_fft)
        y[1, : n // 2] = 0.0
        got = sb.feature.melspectrogram_librosa(y, cfg)
        want = o.librosa_melspectrogram(y, sr, n_fft, hop, n_mels, fmin, fmax)
        assert got.shape == want.shape == (3, n_mels, 1 + n // hop)
        assert_spec_close(got, want, what=f"librosa mel {n_fft}/{hop}/{n_mels}")
        assert_spec_close(sb.feature.melspectrogram_librosa(y[2], cfg), want[2], what="1-D input")
        gdb = sb.feature.melspectrogram_librosa(y, cfg, db=True)
        wdb = o.librosa_melspectrogram(y, sr, n_fft, hop, n_mels, fmin, fmax, db=True)
        loud = want > 1e-6 * want.max()
        assert_log_close(gdb[loud], wdb[loud], atol=2e-3, what="power_to_db")   # 1e-4 relative = 4.3e-4 dB
        assert (gdb >= -100.001).all()
    cfg = sb.feature.librosa_config(16000, 400, 160, 80)
    with pytest.raises(sb.SpeechsauceError):
        sb.feature.melspectrogram_librosa(np.zeros(200, np.float32), cfg)          # reflect padding needs > n_fft/2 samples
    with pytest.raises(sb.SpeechsauceError):
        sb.feature.mel_spectrogram2(np.zeros((1, 4000), np.float32), cfg)          # reference-semantics entry point
    with pytest.raises(sb.SpeechsauceError):
        sb.feature.melspectrogram_librosa(np.zeros(4000, np.float32), sb.SpeechConfig(**CONFIGS["c4"]))


@pytest.mark.parametrize("hop,n_mels,fmin,fmax,n,rows", [(160, 80, 0.0, 8000.0, 480000, 2), (160, 128, 20.0, 7600.0, 16001, 3),
                                                         (100, 40, 0.0, None, 2077, 4), (400, 80, 0.0, None, 201, 2),
                                                         (57, 23, 300.0, 3400.0, 4000, 5)])
def test_librosa_mode_n_fft_400_register_fft_kernel(hop, n_mels, fmin, fmax, n, rows):
    """n_fft = 400 (Whisper's front end) runs on the register-FFT kernel with a full-spectrum |X|^2 tile: staged
    interior rounds (hop 160), per-sample reflect edges, odd hops, the shortest legal row, 128 filters; against
    the oracle and against the generic kernel (SS_FLAG_FORCE_GENERIC)."""
    cfg = sb.feature.librosa_config(16000, 400, hop, n_mels, fmin, fmax)
    y = o.synth_signal(rows, n, 0x5EED0400 + hop)
    y[rows - 1, n // 3:] = 0.0
    got = sb.feature.melspectrogram_librosa(y, cfg)
    assert sb.lib.ss_last_kernel_name() == b"k_stft400m<librosa>"
    want = o.librosa_melspectrogram(y, 16000, 400, hop, n_mels, fmin, fmax)
    assert got.shape == want.shape == (rows, n_mels, 1 + n // hop)
    assert_spec_close(got, want, what=f"librosa mel 400/{hop}/{n_mels}")
    gdb = sb.feature.melspectrogram_librosa(y, cfg, db=True)
    wdb = o.librosa_melspectrogram(y, 16000, 400, hop, n_mels, fmin, fmax, db=True)
    loud = want > 1e-6 * want.max()
    assert_log_close(gdb[loud], wdb[loud], atol=2e-3, what="power_to_db")
    assert (gdb >= -100.001).all()
    gen = sb.feature.librosa_config(16000, 400, hop, n_mels, fmin, fmax, flags=_lib.SS_FLAG_FORCE_GENERIC)
    ref = sb.feature.melspectrogram_librosa(y, gen)
    assert sb.lib.ss_last_kernel_name() == b"k_frames_generic"
    assert_spec_close(got, ref, what="register-FFT kernel vs generic kernel")


def test_cmvn():
    rng = np.random.default_rng(3)
    v = rng.uniform(0, 1, (50, 100)).astype(np.float32)  # the reference's test (lib.rs:71-91)
    y = sb.cmvn(v, True)
    assert y.shape == v.shape
    np.testing.assert_allclose(y.mean(0), 0, atol=1e-6)
    np.testing.assert_allclose(y.std(0), 1, atol=1e-5)
    np.testing.assert_allclose(y, o.cmvn(v, True), atol=2e-5)
    np.testing.assert_allclose(sb.cmvn(v, False), o.cmvn(v, False), atol=1e-6)
    big = rng.standard_normal((100_000, 13)).astype(np.float32) * 3 + 5
    np.testing.assert_allclose(sb.cmvn(big, True), o.cmvn(big, True, np.float64), atol=1e-4)
    batch = rng.standard_normal((40, 97, 13)).astype(np.float32)
    got = sb.cmvn(batch, True)
    for i in (0, 13, 39):
        np.testing.assert_allclose(got[i], o.cmvn(batch[i], True), atol=2e-5)
    # many small matrices: the fused one-warp-per-matrix kernel (13 columns: two row groups; 100 columns: one)
    # ... and the register-resident kernel for the usual cepstral counts (12, 13, 20 columns; up to 64 elements per lane)
    for shape in [(300, 97, 13), (200, 50, 100), (129, 1, 7), (70, 31, 16), (64, 300, 17), (50, 60, 12), (33, 40, 20), (129, 1, 13),
                  (77, 128, 13), (77, 129, 13), (9, 2, 20)]:
        xb = (rng.standard_normal(shape) * 2 + 1).astype(np.float32)
        for vn in (True, False):
            got = sb.cmvn(xb, vn)
            for i in (0, shape[0] // 2, shape[0] - 1):
                np.testing.assert_allclose(got[i], o.cmvn(xb[i], vn), atol=2e-5, err_msg=f"{shape} {vn} {i}")
    # several long matrices: the cooperative L2 pipeline, shifted-sum statistics (a large offset must not hurt the variance)
    long_ = (rng.standard_normal((5, 300_000, 13)) * 0.5 + 40).astype(np.float32)
    got = sb.cmvn(long_, True)
    for i in (0, 4):
        np.testing.assert_allclose(got[i], o.cmvn(long_[i], True, np.float64), atol=2e-4)
    with pytest.raises(sb.SpeechsauceError):
        sb.cmvn(np.zeros((0, 4), np.float32))


# ----------------------------------------------------------------------------------- STFT family


@pytest.mark.parametrize("cname,n_clips,ttd", [("c3", 6, 22050 * 2 + 101), ("c4", 5, 16000 * 2), ("c4", 3, 4001)])
def test_stft_and_mel_match_oracle(cname, n_clips, ttd):
    kw = CONFIGS[cname]
    cfg, oc = sb.SpeechConfig(**kw), o.SpeechConfig(**kw)
    x = o.synth_signal(n_clips, ttd, 31)
    X = sb.functions.stft2(x, cfg)
    if ttd >= 2048:  # the complex STFT comes from the register-FFT kernels too
        assert sb.lib.ss_last_kernel_name() == {"c4": b"k_stft400m", "c3": b"k_stft2048"}[cname]
        Xg = sb.functions.stft2(x, sb.SpeechConfig(**kw, flags=GENERIC))
        assert_spec_close(X.real, Xg.real, what="stft re vs generic")
        assert_spec_close(X.imag, Xg.imag, what="stft im vs generic")
    Xo = o.stft2(x, oc)
    assert X.shape == Xo.shape
    n_pad = oc.window_size // oc.frame_size - 1
    assert not X[:, X.shape[1] - n_pad:].any()            # trailing zero rows (Q13)
    assert_spec_close(X.real, Xo.real, what=f"{cname} stft re")
    assert_spec_close(X.imag, Xo.imag, what=f"{cname} stft im")
    np.testing.assert_array_equal(cfg.analysis_mem, oc.analysis_mem)   # state carried like the reference
    # second call continues from the stored memory (never reset, functions.rs:89-93)
    X2, Xo2 = sb.functions.stft2(x, cfg), o.stft2(x, oc)
    assert_spec_close(X2.real, Xo2.real, what=f"{cname} stft 2nd call")
    cfg2, oc2 = sb.SpeechConfig(**kw), o.SpeechConfig(**kw)
    M = sb.feature.mel_spectrogram2(x, cfg2)
    Mo = o.mel_spectrogram2(x, oc2)
    assert M.shape == Mo.shape == (n_clips, oc.num_filters, o.stft_num_frames(ttd, oc.frame_size))
    assert_spec_close(M, Mo, what=f"{cname} mel_spectrogram2")
    assert not M[:, :, M.shape[2] - n_pad:].any()
    cfg3, oc3 = sb.SpeechConfig(**kw), o.SpeechConfig(**kw)
    assert_log_close(sb.feature.mel_spectrogram2(x, cfg3, log=True), o.log_mel_spectrogram2(x, oc3), what=f"{cname} log-mel")
    cfg4, oc4 = sb.SpeechConfig(**kw), o.SpeechConfig(**kw)
    assert_spec_close(sb.feature.mel_spectrogram1(x[0], cfg4), o.mel_spectrogram1(x[0], oc4), what="mel_spectrogram1")
    assert_spec_close(sb.functions.stft1(x[1], cfg4).real, o.stft1(x[1], oc4).real, what="stft1 after state")


@pytest.mark.parametrize("cname,shapes", [("c4", [(5, 32000), (3, 4001), (7, 161), (2, 100), (9, 3200), (1, 48017)]),
                                          ("c3", [(3, 22050), (2, 5001), (4, 2048), (2, 700), (1, 110250), (37, 16896)])])
def test_stft_fast_path_matches_generic_and_oracle(cname, shapes):
    """the size-specialised STFT kernels (fft_points = 400 / 2048) against the generic kernel and the oracle,
    including ragged lengths, rows shorter than the overlap memory and row-to-row carry"""
    kw = CONFIGS[cname]
    for n, ttd in shapes:
        x = o.synth_signal(n, ttd, 1000 + ttd)
        for log in (False, True):
            fast, gen, oc = sb.SpeechConfig(**kw), sb.SpeechConfig(**kw, flags=GENERIC), o.SpeechConfig(**kw)
            assert fast.info.fast_stft == 1 and gen.info.fast_stft == 0
            a = sb.feature.mel_spectrogram2(x, fast, log=log)
            b = sb.feature.mel_spectrogram2(x, gen, log=log)
            want = o.log_mel_spectrogram2(x, oc) if log else o.mel_spectrogram2(x, oc)
            assert a.shape == b.shape == want.shape
            if log:
                assert_log_close(a, want, what=f"{cname} fast log-mel {n}x{ttd}")
                assert_log_close(a, b, what=f"{cname} fast vs generic log-mel {n}x{ttd}")
            else:
                assert_spec_close(a, want, what=f"{cname} fast mel {n}x{ttd}")
                assert_spec_close(a, b, what=f"{cname} fast vs generic mel {n}x{ttd}")
            # the plan's overlap memory is carried the same way on both paths
            a2 = sb.feature.mel_spectrogram2(x, fast, log=log)
            want2 = o.log_mel_spectrogram2(x, oc) if log else o.mel_spectrogram2(x, oc)
            (assert_log_close if log else assert_spec_close)(a2, want2, what=f"{cname} fast 2nd call {n}x{ttd}")


W512 = [dict(sample_rate=16000, fft_points=512, frame_length=256.5 / 16000, frame_stride=0.01, num_cepstral=13, num_filters=64,
             low_frequency=0.0, high_frequency=8000.0),           # H = 256 | W: the overlap memory only reaches dropped frames
        dict(sample_rate=16000, fft_points=512, frame_length=160.5 / 16000, frame_stride=0.01, num_cepstral=13, num_filters=40,
             low_frequency=50.0, high_frequency=7000.0),          # H = 160: every row's first frames read the previous row's tail
        dict(sample_rate=22050, fft_points=512, frame_length=101.5 / 22050, frame_stride=0.01, num_cepstral=13, num_filters=23,
             low_frequency=0.0, high_frequency=11025.0),          # odd hop: the per-sample path for every frame
        # fft_points = 1024: the same kernel template with 32 points per lane (two 16-point sub-FFTs per lane in the second stage)
        dict(sample_rate=16000, fft_points=1024, frame_length=256.5 / 16000, frame_stride=0.01, num_cepstral=13, num_filters=64,
             low_frequency=0.0, high_frequency=8000.0),
        dict(sample_rate=22050, fft_points=1024, frame_length=441.5 / 22050, frame_stride=0.01, num_cepstral=13, num_filters=96,
             low_frequency=20.0, high_frequency=11025.0),         # odd hop
        # fft_points = 256 (8 kHz speech): 8 points per lane, half of a frame's lanes own rows in the second stage
        dict(sample_rate=8000, fft_points=256, frame_length=80.5 / 8000, frame_stride=0.01, num_cepstral=13, num_filters=26,
             low_frequency=0.0, high_frequency=4000.0),
        dict(sample_rate=8000, fft_points=256, frame_length=77.5 / 8000, frame_stride=0.01, num_cepstral=13, num_filters=20,
             low_frequency=100.0, high_frequency=3800.0)]         # odd hop


@pytest.mark.parametrize("kw", W512)
def test_stft512_register_fft_path_matches_generic_and_oracle(kw):
    """fft_points = 512: k_stft512 (+ k_mel_from_power) against the any-length kernel and the oracle -- complex STFT, mel and
    log-mel, ragged lengths, rows shorter than the overlap memory, row-to-row and call-to-call carry, odd frame counts"""
    kname = {256: b"k_stft256", 512: b"k_stft512", 1024: b"k_stft1024"}[kw["fft_points"]]
    for n, ttd in [(5, 16000), (3, 4001), (7, 257), (2, 100), (1, 48017), (33, 2048)]:
        x = o.synth_signal(n, ttd, 2000 + ttd)
        fast, gen, oc = sb.SpeechConfig(**kw), sb.SpeechConfig(**kw, flags=GENERIC), o.SpeechConfig(**kw)
        assert fast.info.fast_stft == 1 and gen.info.fast_stft == 0
        for call in range(2):  # the second call starts from the memory the first one left
            a = sb.functions.stft2(x, fast)
            assert sb.lib.ss_last_kernel_name() == kname
            b, want = sb.functions.stft2(x, gen), o.stft2(x, oc)
            assert sb.lib.ss_last_kernel_name() == b"k_frames_generic"
            pk = np.abs(want).max()
            assert np.abs(a - want).max() <= 2e-6 * pk, f"stft {n}x{ttd} call {call}"
            assert np.abs(a - b).max() <= 2e-6 * pk
            assert (a[:, a.shape[1] - fast.n_pad:] == 0).all()
        for log in (False, True):
            fast, gen, oc = sb.SpeechConfig(**kw), sb.SpeechConfig(**kw, flags=GENERIC), o.SpeechConfig(**kw)
            a = sb.feature.mel_spectrogram2(x, fast, log=log)
            assert sb.lib.ss_last_kernel_name() == kname
            b = sb.feature.mel_spectrogram2(x, gen, log=log)
            want = o.log_mel_spectrogram2(x, oc) if log else o.mel_spectrogram2(x, oc)
            if log:
                loud = want > np.log(1e-6 * np.exp(want).max())
                assert_log_close(a[loud], want[loud], what=f"512 log-mel {n}x{ttd}")
                assert_spec_close(np.exp(a.astype(np.float64)), np.exp(want.astype(np.float64)), what=f"512 exp(log-mel) {n}x{ttd}")
            else:
                assert_spec_close(a, want, what=f"512 mel {n}x{ttd}")
                assert_spec_close(a, b, what=f"512 fast vs generic mel {n}x{ttd}")


@pytest.mark.parametrize("n_fft", [256, 512, 1024])
def test_librosa_mode_n_fft_512_1024_register_fft_kernel(n_fft):
    kname = {256: b"k_stft256", 512: b"k_stft512", 1024: b"k_stft1024"}[n_fft]
    for sr, hop, n_mels, n, rows in [(16000, 160, 80, 16000 * 2 + 5, 3), (16000, 128, 40, 4000, 2), (22050, 512, 64, 9001, 4), (8000, 77, 23, 600, 1)]:
        if hop > n_fft:  # the plan needs hop <= n_fft
            continue
        cfg = sb.feature.librosa_config(sr, n_fft, hop, n_mels, 0.0, None)
        assert cfg.info.fast_stft == 1
        y = o.synth_signal(rows, n, 0x5EED0512 + hop)
        got = sb.feature.melspectrogram_librosa(y, cfg)
        assert sb.lib.ss_last_kernel_name() == kname
        want = o.librosa_melspectrogram(y, sr, n_fft, hop, n_mels, 0.0, None)
        assert got.shape == want.shape == (rows, n_mels, 1 + n // hop)
        assert_spec_close(got, want, what=f"librosa {n_fft} mel hop {hop}")
        gdb = sb.feature.melspectrogram_librosa(y, cfg, db=True)
        wdb = o.librosa_melspectrogram(y, sr, n_fft, hop, n_mels, 0.0, None, db=True)
        loud = want > 1e-6 * want.max()
        assert_log_close(gdb[loud], wdb[loud], atol=2e-3, what=f"librosa {n_fft} power_to_db")


def test_power_scratch_slabs_of_the_two_kernel_paths(monkeypatch):
    """The |X|^2 scratch of k_stft512 / k_stft1024 / k_stft2048<2> + k_mel_from_power is processed in slabs of rows: with
    a 64 KB slab a small batch crosses many slab boundaries and must give the bits of the one-slab run."""
    cases = [(sb.feature.librosa_config(16000, 512, 160, 80, 0.0, None), 9, 3000),
             (sb.feature.librosa_config(22050, 1024, 256, 64, 0.0, None), 7, 5000),
             (sb.feature.librosa_config(22050, 2048, 512, 128, 0.0, None), 11, 9000)]
    for cfg, rows, n in cases:
        y = o.synth_signal(rows, n, 0x51AB + n)
        monkeypatch.delenv("SS_POWER_SLAB_KB", raising=False)
        whole = sb.feature.melspectrogram_librosa(y, cfg)
        monkeypatch.setenv("SS_POWER_SLAB_KB", "64")
        slabs = sb.feature.melspectrogram_librosa(y, cfg)
        assert np.array_equal(whole, slabs)
    kw = dict(sample_rate=16000, fft_points=1024, frame_length=256.5 / 16000, frame_stride=0.01, num_cepstral=13, num_filters=64,
              low_frequency=0.0, high_frequency=8000.0)
    x = o.synth_signal(9, 4001, 0x51AC)
    monkeypatch.delenv("SS_POWER_SLAB_KB", raising=False)
    whole = sb.feature.mel_spectrogram2(x, sb.SpeechConfig(**kw), log=True)
    monkeypatch.setenv("SS_POWER_SLAB_KB", "64")
    assert np.array_equal(whole, sb.feature.mel_spectrogram2(x, sb.SpeechConfig(**kw), log=True))


def test_librosa_mode_n_fft_2048_register_fft_kernel():
    """librosa's default n_fft: k_stft2048<power> -> |X|^2 rows of all 1025 bins -> k_mel_from_power over the Slaney bank."""
    for sr, hop, n_mels, n, rows in [(22050, 512, 128, 22050 * 2 + 3, 3), (16000, 160, 80, 16000, 2), (44100, 1024, 64, 30001, 2),
                                     (16000, 250, 40, 1100, 4), (22050, 512, 128, 2048, 1)]:
        cfg = sb.feature.librosa_config(sr, 2048, hop, n_mels, 0.0, None)
        assert cfg.info.fast_stft == 1
        y = o.synth_signal(rows, n, 0x5EED2048 + hop)
        got = sb.feature.melspectrogram_librosa(y, cfg)
        assert sb.lib.ss_last_kernel_name() == b"k_stft2048"
        want = o.librosa_melspectrogram(y, sr, 2048, hop, n_mels, 0.0, None)
        assert got.shape == want.shape == (rows, n_mels, 1 + n // hop)
        assert_spec_close(got, want, what=f"librosa 2048 mel hop {hop}")
        gen = sb.feature.librosa_config(sr, 2048, hop, n_mels, 0.0, None, flags=_lib.SS_FLAG_FORCE_GENERIC)
        assert_spec_close(got, sb.feature.melspectrogram_librosa(y, gen), what=f"librosa 2048 fast vs generic hop {hop}")
        gdb = sb.feature.melspectrogram_librosa(y, cfg, db=True)
        wdb = o.librosa_melspectrogram(y, sr, 2048, hop, n_mels, 0.0, None, db=True)
        loud = want > 1e-6 * want.max()
        assert_log_close(gdb[loud], wdb[loud], atol=2e-3, what="librosa 2048 power_to_db")


def test_stft_explicit_memory_equals_plan_memory_and_sharding_halo():
    kw = CONFIGS["c4"]
    x = o.synth_signal(8, 3200, 77)
    whole = sb.functions.stft2(x, sb.SpeechConfig(**kw))
    # process rows [0,3) and [3,8) separately, carrying the overlap memory by hand
    cfg = sb.SpeechConfig(**kw)
    mem = np.zeros(400 - 160, np.float32)
    a = sb.functions.stft2(x[:3], cfg, analysis_mem=mem)
    b = sb.functions.stft2(x[3:], cfg, analysis_mem=mem)
    np.testing.assert_array_equal(np.concatenate([a, b]), whole)
    assert not cfg.analysis_mem.any()  # the plan's own memory was not touched


def test_stft_impulse_closed_form():
    W, H = 400, 160
    cfg = sb.SpeechConfig(16000, W, 0.01, 0.01, 13, 80, 0.0, 8000.0)
    ttd, p = 10 * H + 37, 3 * H + 11
    x = np.zeros((1, ttd), np.float32)
    x[0, p] = 1.0
    X = sb.functions.stft2(x, cfg)
    win = cfg.window
    k = np.arange(W // 2 + 1)
    for t in range(X.shape[1] - 1):
        off = p - ((t + 2) * H - W)
        want = win[off] * cfg.wnorm * np.exp(-2j * np.pi * k * off / W) if 0 <= off < W else 0 * k
        np.testing.assert_allclose(X[0, t], want, atol=2e-8)


def test_stft_errors():
    dflt = sb.SpeechConfig(16000)  # H = 320 > W/2: the reference underflows (Q13)
    with pytest.raises(sb.SpeechsauceError) as ei:
        sb.functions.stft2(np.zeros((1, 4000), np.float32), dflt)
    assert ei.value.code == _lib.SS_ERR_INVALID_ARG


def test_cpp_mirror_runs_reference_unit_tests():
    """builds mfcc-rust_b200/cpp/example.cpp (the reference's lib.rs:41-134 tests on the C++ mirror
    of the crate API) against the shipped library and runs it on this GPU"""
    import subprocess
    from pathlib import Path
    pkg = Path(sb.LIB_PATH).parent.parent
    exe = Path("/tmp/ss_cpp_example")
    subprocess.check_call(["g++", "-O2", "-std=c++17", f"-I{pkg.parent / 'include'}", f"-I{pkg / 'cpp'}",
                           str(pkg / "cpp" / "example.cpp"), "-o", str(exe), f"-L{pkg / 'lib'}", "-lspeechsauce_cuda",
                           f"-Wl,-rpath,{pkg / 'lib'}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "all reference-shaped tests passed" in out.stdout


def test_multi_device_plan_matches_single_device():
    """clips shard over a plan's devices with no collective; results must not depend on the split"""
    n = sb.lib.ss_device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    x = o.synth_signal(37, 16000, 77)
    one = sb.feature.features(x, sb.SpeechConfig(**CONFIGS["c1"], devices=[0]), lmfe=True, mfcc=True, cmvn_mode=2)
    many = sb.feature.features(x, sb.SpeechConfig(**CONFIGS["c1"], devices=list(range(n))), lmfe=True, mfcc=True, cmvn_mode=2)
    np.testing.assert_array_equal(one["mfcc"], many["mfcc"])
    np.testing.assert_array_equal(one["lmfe"], many["lmfe"])
    y = o.synth_signal(9, 8000, 78)
    a = sb.functions.stft2(y, sb.SpeechConfig(**CONFIGS["c4"], devices=[0]))
    b = sb.functions.stft2(y, sb.SpeechConfig(**CONFIGS["c4"], devices=list(range(n))))
    np.testing.assert_array_equal(a, b)


# ------------------------------------------------------ "next" rows: deltas, cmvnw


def test_derivative_features_bit_exact():
    rng = np.random.default_rng(11)
    for shape in [(97, 13), (1, 40), (359, 5), (1000, 1), (64, 3)]:
        f = rng.standard_normal(shape).astype(np.float32)
        for dw in (1, 2, 3):
            np.testing.assert_array_equal(sb.processing.derivative_extraction(f, dw), o.derivative_extraction(f, dw))
        np.testing.assert_array_equal(sb.feature.extract_derivative_feature(f), o.extract_derivative_feature(f))
    with pytest.raises(sb.SpeechsauceError):
        sb.processing.derivative_extraction(np.zeros((4, 4), np.float32), 0)


def test_cmvn_long_matrices_cooperative_pipeline():
    """Matrices of 3 .. 40 MB: the co-resident worker / merger pipeline (k_cmvn_coop): more matrices than the lag and
    the merger count, both modes, a hostile first row (the pivot of block 0), and the same bits as a one-matrix call.
    Smaller ones: the partial / merge / apply kernels under the same conditions."""
    rng = np.random.default_rng(11)
    # the last two shapes (1.6 MB matrices) stay on the three-pass kernels: same pivot, same tolerance
    for shape in [(11, 70_000, 13), (3, 120_000, 40), (4, 30_000, 13), (3, 5_000, 80)]:
        xb = (rng.standard_normal(shape) * 2 + 1).astype(np.float32)
        xb[:, 0, :] = 1e4
        for vn in (True, False):
            got = sb.cmvn(xb, vn)
            for i in range(shape[0]):
                np.testing.assert_allclose(got[i], o.cmvn(xb[i], vn, np.float64), atol=2e-4, err_msg=f"{shape} {vn} {i}")
            one = sb.cmvn(xb[shape[0] - 1], vn)
            assert np.array_equal(one, got[shape[0] - 1])


def test_cmvnw_matches_oracle():
    rng = np.random.default_rng(12)
    for rows, cols, win in [(1000, 13, 301), (50, 13, 301), (97, 40, 51), (20000, 13, 301), (7, 3, 1), (300, 1, 299)]:
        x = (rng.standard_normal((rows, cols)) * 3 + 5).astype(np.float32)
        for vn in (False, True):
            got = sb.processing.cmvnw(x, win, vn)
            np.testing.assert_allclose(got, o.cmvnw(x, win, vn), atol=2e-5, rtol=2e-5)
            if rows <= 1000:
                np.testing.assert_allclose(got, o.cmvnw_direct(x, win, vn), atol=5e-5, rtol=5e-5)
    with pytest.raises(sb.SpeechsauceError, match="odd"):
        sb.processing.cmvnw(np.zeros((10, 3), np.float32), 300, False)   # processing.rs:327


# ---------------------------------------------------------------- randomised configuration sweep
# SS_SWEEP_BLOCKS=n widens every sweep to n blocks of 10-12 parameter sets (defaults 4 / 3 / 2 / 2; 100 / 100 / 20 / 20 =
# 2,800 parameter sets ran clean on a B200)
import os

_SWEEP = int(os.environ.get("SS_SWEEP_BLOCKS", "0"))


def _random_speechpy_config(rng):
    sr = int(rng.choice([8000, 11025, 16000, 22050, 44100, 48000]))
    n_fft = int(rng.choice([64, 128, 200, 256, 384, 400, 480, 512, 512, 512, 512, 640, 1000, 1024, 1536, 2048, 4096]))
    L = int(rng.integers(max(8, n_fft // 4), n_fft + 1))
    S = int(rng.integers(max(1, L // 5), L + 1))
    nf = int(rng.integers(4, 81))
    nc = int(rng.integers(1, min(nf, 24) + 1))
    lo = float(rng.choice([0.0, 0.0, 64.0, 133.33]))
    hi = float(rng.choice([sr / 2, sr / 2, 0.9 * sr / 2, sr / 3]))
    # + 0.25 keeps round(frame_length * sr) (processing.rs:77-78) away from the .5 ties
    return dict(sample_rate=sr, fft_points=n_fft, frame_length=(L + 0.25) / sr, frame_stride=(S + 0.25) / sr, num_cepstral=nc,
                num_filters=nf, low_frequency=lo, high_frequency=hi, dc_elimination=bool(rng.integers(0, 2)))


@pytest.mark.parametrize("block", range(_SWEEP or 4))
def test_random_speechpy_configs_match_oracle(block):
    """48 seeded random parameter sets (sample rate, FFT length incl. mixed-radix ones, frame length / stride,
    filter and cepstrum counts, band edges, dc_elimination) x ragged clip lengths: whatever kernel the plan picks
    (k_fast512c inside its limits, the generic kernel otherwise) against the oracle, every output of the fused
    call, with and without the fused CMVN."""
    rng = np.random.Generator(np.random.Philox(0x5EED0600 + block))
    picked = set()
    for i in range(12):
        kw = _random_speechpy_config(rng)
        cfg, oc = sb.SpeechConfig(**kw), o.SpeechConfig(**kw)
        L, S = oc.frame_sample_length, oc.frame_step
        assert (cfg.frame_sample_length, cfg.frame_step) == (L, S)
        n = L + S * int(rng.integers(1, 40)) + int(rng.integers(0, S))
        n_clips = int(rng.integers(1, 6))
        x = o.synth_signal(n_clips, n, 0x5EED0700 + 97 * block + i)
        mode = int(rng.integers(0, 3))
        r = sb.feature.features(x, cfg, mfe=True, energy=True, lmfe=True, ssc=True, mfcc=True, cmvn_mode=mode)
        picked.add(sb.lib.ss_last_kernel_name())
        what = f"block {block} config {i} {kw} n={n}"
        for c in range(n_clips):
            f, e = o.mfe(x[c], oc)
            assert r["mfe"][c].shape == f.shape, what
            assert_spec_close(r["mfe"][c], f, what=what + " mfe")
            assert_spec_close(r["energy"][c], e, what=what + " energy")
            # a one-bin filter on a bin where |X| happens to be 1e-4 of its usual size holds the f32 FFT's own
            # rounding (about 1 in 10^6 entries on this noise): those entries are checked in the linear domain
            loud = f > 1e-3 * f.max()
            assert_log_close(r["lmfe"][c][loud], np.log(f[loud]), what=what + " lmfe")
            assert_spec_close(np.exp(r["lmfe"][c].astype(np.float64)), f, what=what + " exp(lmfe)")
            raw = o.mfcc(x[c], oc)
            m = o.cmvn(raw, mode == 2) if mode else raw
            # variance normalisation divides by std + 2^-30: a column that is (nearly) constant over a short clip
            # turns the f32 rounding of the cepstra into O(1) differences on both sides
            ok = raw.std(axis=0) > 0.05 if mode == 2 else np.ones(m.shape[1], bool)
            rows = loud.all(axis=1)       # frames without such an entry
            got = r["mfcc"][c][rows]
            if not kw["dc_elimination"] and mode != 2:
                # c0 = 2 sum_f ln(mel_f) is unscaled except for element [0,0] (Q9) and reaches 1e3..2e3: its f32
                # rounding alone is 1e-4 per term, so the abs 1e-3 bound is applied to columns 1.. only
                assert_log_close(got[:, 0], m[rows, 0], atol=4e-3, what=what + f" mfcc c0 cmvn={mode}")
                ok[0] = False
            assert_log_close(got[:, ok], m[rows][:, ok], what=what + f" mfcc cmvn={mode}")
            assert_spec_close(r["ssc"][c], o.ssc(x[c], oc), what=what + " ssc")
    assert len(picked) >= 1


def _random_stft_config(rng):
    sr = int(rng.choice([8000, 16000, 22050, 44100]))
    W = int(rng.choice([64, 128, 200, 256, 400, 400, 480, 512, 1000, 1024, 2048, 2048]))
    H = int(rng.integers(max(1, W // 16), W // 2 + 1))          # 0 < H <= W / 2 (functions.rs:96)
    nf = int(rng.integers(4, 129))
    lo = float(rng.choice([0.0, 0.0, 50.0]))
    hi = float(rng.choice([sr / 2, sr / 2, 0.8 * sr / 2]))
    return dict(sample_rate=sr, fft_points=W, frame_length=(H + 0.5) / sr, frame_stride=0.01, num_cepstral=min(13, nf),
                num_filters=nf, low_frequency=lo, high_frequency=hi)


@pytest.mark.parametrize("block", range(_SWEEP or 3))
def test_random_stft_configs_match_oracle(block):
    """36 seeded random STFT parameter sets (window 64..2048 incl. mixed-radix, hop anywhere in (0, W/2] so that
    H does not divide W in most of them, 4..128 filters) x ragged row lengths: complex STFT, mel spectrogram and
    its log against the oracle, including the row-to-row and call-to-call overlap memory (Q13)."""
    rng = np.random.Generator(np.random.Philox(0x5EED0800 + block))
    for i in range(12):
        kw = _random_stft_config(rng)
        oc = o.SpeechConfig(**kw)
        W, H = oc.window_size, oc.frame_size
        ttd = int(rng.integers(1, 30)) * H + int(rng.integers(0, H)) + (W if rng.integers(0, 2) else 0)
        rows = int(rng.integers(1, 5))
        x = o.synth_signal(rows, ttd, 0x5EED0900 + 97 * block + i)
        what = f"block {block} config {i} {kw} rows={rows} ttd={ttd}"
        cfg = sb.SpeechConfig(**kw)
        assert cfg.frame_size == H, what
        for call in range(2):                                  # the second call starts from the carried memory
            X, Xo = sb.functions.stft2(x, cfg), o.stft2(x, oc)
            assert X.shape == Xo.shape, what
            assert_spec_close(X.real, Xo.real, what=what + f" stft re call {call}")
            assert_spec_close(X.imag, Xo.imag, what=what + f" stft im call {call}")
            np.testing.assert_array_equal(cfg.analysis_mem, oc.analysis_mem)
        cfg2, oc2 = sb.SpeechConfig(**kw), o.SpeechConfig(**kw)
        M, Mo = sb.feature.mel_spectrogram2(x, cfg2), o.mel_spectrogram2(x, oc2)
        assert M.shape == Mo.shape, what
        assert_spec_close(M, Mo, what=what + " mel")
        cfg3, oc3 = sb.SpeechConfig(**kw), o.SpeechConfig(**kw)
        Ml = sb.feature.mel_spectrogram2(x, cfg3, log=True)
        Mo3 = o.mel_spectrogram2(x, oc3)
        loud = Mo3 > 1e-6 * Mo3.max()
        assert_log_close(Ml[Mo3 == 0], np.full((Mo3 == 0).sum(), np.log(o.EPS)), atol=1e-5, what=what + " ln(zero_handling(0))")
        assert_log_close(Ml[loud], np.log(Mo3[loud].astype(np.float64)), what=what + " log-mel")


@pytest.mark.parametrize("block", range(_SWEEP or 2))
def test_random_librosa_configs_match_oracle(block):
    """seeded random librosa-mode parameter sets (n_fft incl. 400 / 2048 and mixed-radix ones, any hop <= n_fft,
    Slaney banks of 8..128 filters with random band edges) x ragged lengths down to the shortest the reflect
    padding allows: mel power spectrogram and power_to_db against the oracle (pinned to torchaudio)."""
    rng = np.random.Generator(np.random.Philox(0x5EED0B00 + block))
    for i in range(10):
        sr = int(rng.choice([8000, 16000, 22050, 44100]))
        n_fft = int(rng.choice([128, 256, 400, 400, 480, 512, 1000, 1024, 2048]))
        hop = int(rng.integers(max(1, n_fft // 16), n_fft + 1))
        n_mels = int(rng.integers(8, 129))
        fmin = float(rng.choice([0.0, 0.0, 30.0, 125.5]))
        fmax = [None, 0.45 * sr, 0.3 * sr][int(rng.integers(0, 3))]
        n = n_fft // 2 + 1 + int(rng.integers(0, 40)) * hop + int(rng.integers(0, hop))
        rows = int(rng.integers(1, 5))
        what = f"block {block} config {i} sr={sr} n_fft={n_fft} hop={hop} n_mels={n_mels} fmin={fmin} fmax={fmax} n={n} rows={rows}"
        cfg = sb.feature.librosa_config(sr, n_fft, hop, n_mels, fmin, fmax)
        y = o.synth_signal(rows, n, 0x5EED0C00 + 97 * block + i)
        got = sb.feature.melspectrogram_librosa(y, cfg)
        want = o.librosa_melspectrogram(y, sr, n_fft, hop, n_mels, fmin, fmax)
        assert got.shape == want.shape == (rows, n_mels, 1 + n // hop), what
        assert_spec_close(got, want, what=what + " mel")
        gdb = sb.feature.melspectrogram_librosa(y, cfg, db=True)
        wdb = o.librosa_melspectrogram(y, sr, n_fft, hop, n_mels, fmin, fmax, db=True)
        loud = want > 1e-6 * want.max()
        assert_log_close(gdb[loud], wdb[loud], atol=2e-3, what=what + " power_to_db")   # 1e-4 relative = 4.3e-4 dB
        assert (gdb >= -100.001).all(), what


@pytest.mark.parametrize("block", range(_SWEEP or 2))
def test_random_processing_primitives_match_oracle(block):
    """processing::{preemphasis, stack_frames, power_spectrum, log_power_spectrum, cmvn, cmvnw,
    derivative_extraction} on seeded random shapes and parameters: bit-exact where the work is indexing or a
    single rounding (preemphasis, stack_frames, derivatives), within the spectrum / log tolerances otherwise."""
    rng = np.random.Generator(np.random.Philox(0x5EED0D00 + block))
    for i in range(10):
        n = int(rng.integers(2, 200_000))
        x = o.synth_signal(1, n, 0x5EED0E00 + 97 * block + i)[0]
        shift = int(rng.choice([1, 1, 2, -1, -7, n, -(n - 1), int(rng.integers(1, n))]))
        cof = float(rng.choice([0.98, 0.97, 0.5, 1.0]))
        what = f"block {block} case {i} n={n}"
        np.testing.assert_array_equal(sb.preemphasis(x, shift, cof), o.preemphasis(x, shift, cof), err_msg=what + f" preemphasis shift={shift}")
        # stack_frames: L, S by round(sr * length) (processing.rs:77-78), floor / ceil frame counts in f32
        sr = int(rng.choice([8000, 16000, 22050, 44100]))
        L = int(rng.integers(2, min(n, 3000) + 1))
        S = int(rng.integers(1, 2 * L))
        zp = bool(rng.integers(0, 2))
        fl, fs = (L + 0.25) / sr, (S + 0.25) / sr
        try:
            want = o.stack_frames(x, sr, fl, fs, zp)
        except ValueError:
            with pytest.raises(sb.SpeechsauceError):
                sb.processing.stack_frames(x, sr, fl, fs, None, zp)
            want = None
        if want is not None:
            got = sb.processing.stack_frames(x, sr, fl, fs, None, zp)
            np.testing.assert_array_equal(got, want, err_msg=what + f" stack_frames L={L} S={S} zp={zp}")
        # power_spectrum / log_power_spectrum on random frame matrices (frame length <= n_fft)
        nfft = int(rng.choice([64, 200, 256, 400, 512, 1000, 1024, 2048]))
        Lf = int(rng.integers(1, nfft + 1))
        rows = int(rng.integers(1, 300))
        fr = o.synth_signal(rows, Lf, 0x5EED0F00 + 97 * block + i)
        ps, po = sb.processing.power_spectrum(fr, nfft), o.power_spectrum(fr, nfft)
        assert_spec_close(ps, po, what=what + f" power_spectrum rows={rows} L={Lf} n_fft={nfft}")
        for norm in (False, True):
            lg, lo_ = sb.processing.log_power_spectrum(fr, nfft, norm), o.log_power_spectrum(fr, nfft, norm)
            loud = po > 1e-5
            if norm and not loud.reshape(-1)[-1]:
                continue           # the normaliser is the last value of the matrix (Q7): noise in, noise everywhere
            assert_log_close(lg[loud], lo_[loud], what=what + f" log_power_spectrum n_fft={nfft} normalize={norm}")
        # cmvn / cmvnw / derivatives on a random feature matrix
        T, C = int(rng.integers(1, 3000)), int(rng.integers(1, 65))
        scale = float(rng.uniform(0.1, 10.0))
        v = (rng.standard_normal((T, C)) * scale + rng.uniform(-20.0, 20.0)).astype(np.float32)
        # the f32 column mean carries ~1e-6 of the largest magnitude whatever the summation order is; variance
        # normalisation divides that by the (local) standard deviation
        tol = 2e-5 * float(np.abs(v).max()) + 2e-5
        for vn in (False, True):
            if vn and T < 30:
                continue           # std of a few rows: 1 / (std + 2^-30) amplifies the last ulp of the mean
            np.testing.assert_allclose(sb.cmvn(v, vn), o.cmvn(v, vn), atol=tol / (0.5 * scale) if vn else tol, rtol=0,
                                       err_msg=what + f" cmvn {T}x{C} vn={vn}")
        win = 2 * int(rng.integers(0, 200)) + 1
        for vn in (False, True):
            if vn and win < 31:
                continue
            np.testing.assert_allclose(sb.processing.cmvnw(v, win, vn), o.cmvnw(v, win, vn), atol=tol / (0.3 * scale) if vn else tol,
                                       rtol=0, err_msg=what + f" cmvnw {T}x{C} win={win} vn={vn}")
        dw = int(rng.integers(1, 5))
        np.testing.assert_array_equal(sb.processing.derivative_extraction(v, dw), o.derivative_extraction(v, dw),
                                      err_msg=what + f" derivative_extraction {T}x{C} dw={dw}")
        np.testing.assert_array_equal(sb.feature.extract_derivative_feature(v), o.extract_derivative_feature(v))
