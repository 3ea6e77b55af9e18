"""CPU tests of the oracle itself: the reference's own test invariants (lib.rs:41-134), the
closed-form known-answer tests of SURVEY.md Appendix D, and the committed golden fixtures."""
import json
from pathlib import Path

import numpy as np
import pytest
import scipy.fft

import speechsauce_oracle as o

GOLD = Path(__file__).parent / "golden"


def default_cfg(sr=16000):
    return o.SpeechConfig(sr)  # SpeechConfigBuilder::new(sr).build(), config.rs:35-47


def ref_get_num_frames(n, sr, fl, fs):  # lib.rs:31-40
    w = np.float32(np.round(np.float32(fl) * np.float32(sr)))
    s = np.float32(np.round(np.float32(fs) * np.float32(sr)))
    return int(np.ceil((np.float32(n) - w) / s))


@pytest.fixture(scope="module")
def signal():
    return o.synth_signal(1, 1_000_000, 7)[0]  # create_signal(), lib.rs:18-22


def test_preemphasis_shape_and_kat(signal):
    assert o.preemphasis(signal, 1, 0.98).shape == signal.shape  # lib.rs:42-48
    got = o.preemphasis(np.array([1, 2, 3, 4, 5], np.float32), 1, 0.98)
    np.testing.assert_allclose(got, [1 - 4.9, 2 - 0.98, 3 - 1.96, 4 - 2.94, 5 - 3.92], rtol=1e-6)
    with pytest.raises(ValueError):
        o.preemphasis(signal, 0, 0.98)


def test_stack_frames_counts_and_gather(signal):
    fr = o.stack_frames(signal, 16000, 0.02, 0.02, zero_padding=True)  # lib.rs:51-68
    assert fr.shape[0] == ref_get_num_frames(signal.size, 16000, 0.02, 0.02)
    for n, L, S, T in [(16000, 400, 160, 97), (16000, 320, 160, 98), (10**6, 320, 160, 6248), (57_600_000, 400, 160, 359_997)]:
        assert o.num_frames(n, L, S, False) == T
    x = np.arange(16000, dtype=np.float32)
    fr = o.stack_frames(x, 16000, 0.025, 0.01, False)
    assert fr.shape == (97, 400)
    assert (fr == (np.arange(97)[:, None] * 160 + np.arange(400)[None, :])).all()
    lit = o.stack_frames(x, 16000, 0.025, 0.01, False, literal=True)
    assert lit.shape == (97, 400) and not lit.any()  # Q1: nothing is copied for T > 2
    lit2 = o.stack_frames(x[:800], 16000, 0.025, 0.01, False, literal=True)  # T == 2
    assert lit2.shape == (2, 400) and (lit2[0] == x[:400]).all() and (lit2[1] == x[:400]).all()


def test_cmvn_reference_invariants():
    rng = np.random.default_rng(3)
    v = rng.uniform(0, 1, (50, 100)).astype(np.float32)  # lib.rs:71-91
    y = o.cmvn(v, True)
    assert y.shape == v.shape
    np.testing.assert_allclose(y.mean(0), 0, atol=1e-6)
    np.testing.assert_allclose(y.std(0), 1, atol=1e-5)


def test_mfcc_mfe_reference_shapes(signal):
    cfg = default_cfg()
    m = o.mfcc(signal, cfg)  # lib.rs:94-114
    assert not np.isnan(m).any()
    T = ref_get_num_frames(signal.size, 16000, cfg.frame_length, cfg.frame_stride)
    assert m.shape == (T, 13) == (6248, 13)
    f, e = o.mfe(signal, cfg)  # lib.rs:116-134
    assert f.shape == (6248, 40) and e.shape == (6248,)


def test_power_spectrum_closed_form():
    P = o.power_spectrum(np.ones((1, 320), np.float32), 512)
    k = np.arange(1, 257)
    want = np.abs(np.sin(np.pi * 320 * k / 512) / np.sin(np.pi * k / 512)) / 512
    assert P[0, 0] == np.float32(0.625)
    np.testing.assert_allclose(P[0, 1:], want, rtol=2e-5, atol=2e-7)


def test_log_power_spectrum_closed_form_and_last_value_normaliser():
    """processing.rs:194-214: 10 log10 P above 1e-20, else -200; `normalize` subtracts the LAST value
    above the threshold in row-major order (Q7), not the maximum."""
    fr = np.zeros((3, 320), np.float32)
    fr[0] = 1.0                      # Dirichlet kernel: every bin above the threshold
    fr[1, 0] = 1e-3                  # impulse: flat spectrum 1e-3 / 512
    # row 2 stays silent: every bin is -200
    raw = o.log_power_spectrum(fr, 512, normalize=False)
    assert raw.shape == (3, 257) and raw.dtype == np.float32
    assert abs(raw[0, 0] - 10 * np.log10(0.625)) < 1e-5
    np.testing.assert_allclose(raw[1], 10 * np.log10(1e-3 / 512), atol=1e-4)
    assert (raw[2] == -200.0).all()
    nrm = o.log_power_spectrum(fr, 512, normalize=True)
    # the last element above the threshold is the impulse row's last bin, not the maximum (row 0, bin 0)
    np.testing.assert_allclose(nrm, raw - raw[1, 256], atol=1e-5)
    assert nrm.max() > 30.0
    silent = o.log_power_spectrum(np.zeros((2, 64), np.float32), 64, normalize=True)
    assert (silent == -200.0).all()   # mx stays 1e-20


def test_librosa_mode_oracle_is_pinned_by_torchaudio():
    """The librosa-compatible mode has no reference code; its oracle restates
    librosa's documented algorithm and is pinned here against an independent implementation of the
    same definition, torchaudio.transforms.MelSpectrogram(norm="slaney", mel_scale="slaney")."""
    torch = pytest.importorskip("torch")
    torchaudio = pytest.importorskip("torchaudio")
    for sr, n_fft, hop, n_mels, fmin, fmax in [(22050, 2048, 512, 128, 0.0, None), (16000, 400, 160, 80, 0.0, 8000.0),
                                               (16000, 512, 128, 40, 50.0, 7000.0)]:
        y = o.synth_signal(2, sr + 37, 0x5EED0300 + n_fft)
        fm = sr / 2.0 if fmax is None else fmax
        ms = torchaudio.transforms.MelSpectrogram(sample_rate=sr, n_fft=n_fft, hop_length=hop, f_min=fmin, f_max=fm,
                                                  n_mels=n_mels, power=2.0, center=True, pad_mode="reflect",
                                                  norm="slaney", mel_scale="slaney")
        want = ms(torch.from_numpy(y)).numpy()
        got = o.librosa_melspectrogram(y, sr, n_fft, hop, n_mels, fmin, fmax)
        assert got.shape == want.shape == (2, n_mels, 1 + (sr + 37) // hop)
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-6 * want.max())
        fb = torchaudio.functional.melscale_fbanks(n_fft // 2 + 1, fmin, fm, n_mels, sr, norm="slaney", mel_scale="slaney")
        np.testing.assert_allclose(o.librosa_mel_filters(sr, n_fft, n_mels, fmin, fmax), fb.numpy().T, atol=2e-6)
    db = o.librosa_melspectrogram(np.zeros(4000, np.float32), 16000, 400, 160, 80, db=True)
    assert np.abs(db + 100.0).max() < 1e-4     # 10 log10(amin = 1e-10)


def test_librosa_mode_oracle_against_torchaudio_on_random_parameter_sets():
    """40 seeded random (sr, n_fft, hop, n_mels, fmin, fmax, length) as in the GPU sweep.  torchaudio builds its
    Slaney bank in f32 (librosa and the oracle in f64), which shows as a few 1e-6 of the spectrogram's peak where
    117 filters share 241 bins; hence the absolute floor of 1e-5 of the peak on top of rel 1e-4."""
    torch = pytest.importorskip("torch")
    torchaudio = pytest.importorskip("torchaudio")
    rng = np.random.Generator(np.random.Philox(0x5EED0A40))
    for i in range(40):
        sr = int(rng.choice([8000, 16000, 22050, 44100]))
        n_fft = int(rng.choice([128, 256, 400, 480, 512, 1000, 1024, 2048]))
        hop = int(rng.integers(max(1, n_fft // 16), n_fft + 1))
        n_mels = int(rng.integers(8, 129))
        fmin = float(rng.choice([0.0, 0.0, 30.0, 125.5]))
        fmax = [None, 0.45 * sr, 0.3 * sr][int(rng.integers(0, 3))]
        n = n_fft // 2 + 1 + int(rng.integers(0, 40)) * hop + int(rng.integers(0, hop))
        y = o.synth_signal(2, n, 900 + i)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")   # "some mel filterbanks are all zero": intended here
            ms = torchaudio.transforms.MelSpectrogram(sample_rate=sr, n_fft=n_fft, hop_length=hop, f_min=fmin,
                                                      f_max=sr / 2.0 if fmax is None else fmax, n_mels=n_mels, power=2.0,
                                                      center=True, pad_mode="reflect", norm="slaney", mel_scale="slaney")
            want = ms(torch.from_numpy(y)).numpy()
        got = o.librosa_melspectrogram(y, sr, n_fft, hop, n_mels, fmin, fmax)
        assert got.shape == want.shape == (2, n_mels, 1 + n // hop)
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5 * want.max(), err_msg=f"{sr} {n_fft} {hop} {n_mels} {fmin} {fmax} {n}")


def test_filterbank_indices_and_peaks():
    want = [0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 13, 15, 17, 18, 20, 22, 25, 27, 29, 32, 35, 38, 41, 44,
            48, 51, 55, 60, 64, 69, 74, 79, 85, 91, 98, 105, 112, 120, 129]
    idx = o.filterbank_indices(40, 257, 16000, 0.0, 8000.0)
    assert list(idx) == want
    for nf, F, sr, hi, taps, empty, top in [(40, 257, 16000, 8000.0, 210, 0, 128), (128, 1025, 22050, 11025.0, 889, 2, 512),
                                            (80, 201, 16000, 8000.0, 137, 18, 100)]:
        fb = o.filterbanks(nf, F, sr, 0.0, hi)
        assert (fb != 0).sum() == taps
        assert (fb.sum(1) == 0).sum() == empty
        assert np.nonzero(fb.sum(0))[0].max() == top
        ix = o.filterbank_indices(nf, F, sr, 0.0, hi)
        for i in range(nf):
            if fb[i].any():
                assert fb[i, ix[i + 1]] == 1.0  # every non-empty filter peaks at exactly 1 at idx[i+1]
        assert not np.isnan(fb).any()
    with pytest.raises(ValueError):
        o.filterbanks(40, 257, 16000, 0.0, 9000.0)  # feature.rs:47-50


def test_f32_ln_exp_are_the_platform_libm_ones():
    """The reference's f32::ln / f32::exp are the platform libm's logf / expf, and so are the oracle's (numpy's
    float32 loops are 1 ulp off on a few percent of the inputs, which moves a band edge by one FFT bin when
    (F + 1) * hz / sr sits on an integer).  glibc's are within 1 ulp of the correctly rounded value."""
    if o._LOGF is None:
        pytest.skip("no libm to compare with")
    rng = np.random.Generator(np.random.Philox(0x5EED0A01))
    hz = np.concatenate([rng.uniform(0.0, 24000.0, 5000), [0.0, 4000.0, 8000.0, 11025.0, 22050.0, 24000.0]]).astype(np.float32)
    arg = np.float32(1.0) + hz / np.float32(700.0)
    ln = o._ln(arg)
    assert ln.dtype == np.float32 and ln.shape == arg.shape
    exact = np.log(arg.astype(np.float64))
    assert (np.abs(ln - exact) <= np.spacing(np.abs(exact).astype(np.float32))).all()
    q = np.float32(1127.0) * ln / np.float32(1127.0)
    ex = o._exp(q)
    exact = np.exp(q.astype(np.float64))
    assert (np.abs(ex - exact) <= np.spacing(exact.astype(np.float32))).all()
    assert o._ln(np.float64(2.0)).dtype == np.float64          # the f64 oracle keeps numpy's f64 functions
    # the case that exposed it: 44.1 kHz, n_fft 1024, 58 filters up to sr / 2 -> the last edge is bin 257, not 256
    assert o.filterbank_indices(58, 513, 44100, 64.0, 22050.0)[-1] == 257


def test_zero_clip_and_literal_framing():
    cfg = o.SpeechConfig(16000, 512, 0.025, 0.01)
    f, e = o.mfe(np.zeros(16000, np.float32), cfg)
    assert (f == o.EPS).all() and (e == o.EPS).all()
    m = o.mfcc(np.zeros(16000, np.float32), cfg)
    np.testing.assert_allclose(m[:, 0], np.log(o.EPS), rtol=1e-6)
    np.testing.assert_allclose(m[:, 1:], 0, atol=1e-5)
    x = o.synth_signal(1, 16000, 5)[0]
    np.testing.assert_array_equal(o.mfcc(x, cfg, literal_framing=True), m)  # Q1: literal output is input-independent


def test_dct2_convention():
    y = o.dct2(np.full((1, 40), 0.5, np.float32))
    assert abs(y[0, 0] - 80 * 0.5) < 1e-4 and np.abs(y[0, 1:]).max() < 1e-4
    x = np.random.default_rng(0).standard_normal((3, 40)).astype(np.float32)
    n = np.arange(40)
    want = np.stack([2 * (x * np.cos(np.pi * k * (2 * n + 1) / 80)).sum(1) for k in range(40)], 1)
    np.testing.assert_allclose(o.dct2(x), want, atol=2e-5)


def test_mfcc_scale_quirk_without_dc_elimination():
    cfg = o.SpeechConfig(16000, 512, 0.025, 0.01, dc_elimination=False)
    x = o.synth_signal(1, 16000, 9)[0]
    m = o.mfcc(x, cfg)
    lf = np.log(o.mfe(x, cfg)[0]).astype(np.float64)
    n = 97 * 40
    assert abs(m[0, 0] - 2 * lf[0].sum() / np.sqrt(4 * n)) < 1e-4   # only [0,0] gets 1/sqrt(4n)
    assert abs(m[5, 0] - 2 * lf[5].sum()) < 2e-3                    # rest of column 0 is unscaled (Q9)


def test_stft_impulse_and_state():
    for W, H in [(2048, 512), (400, 160)]:
        cfg = o.SpeechConfig(16000, W, H / 16000 + 1e-9, 0.01)
        assert cfg.frame_size == H
        n_pad = W // H - 1
        ttd = 10 * H + 37
        p = 3 * H + 11
        x = np.zeros((1, ttd), np.float32)
        x[0, p] = 1.0
        X = o.stft2(x, cfg)
        Tp = o.stft_num_frames(ttd, H)
        assert X.shape == (1, Tp, W // 2 + 1)
        assert not X[0, Tp - n_pad:].any()  # trailing zero rows
        k = np.arange(W // 2 + 1)
        for t in range(Tp - n_pad):
            j = t + n_pad
            off = p - ((j + 1) * H - W)
            want = (cfg.window[off] * cfg.wnorm * np.exp(-2j * np.pi * k * off / W)) if 0 <= off < W else 0 * k
            np.testing.assert_allclose(X[0, t], want, atol=1e-8)
        # second call on the same config: row 0 changes only when H does not divide W
        y = o.synth_signal(1, 10 * H, 11)  # exact multiple of H: the carried tail is signal, not padding
        cfgA = o.SpeechConfig(16000, W, H / 16000 + 1e-9, 0.01)
        first = o.stft2(y, cfgA)
        second = o.stft2(y, cfgA)
        if W % H == 0:
            np.testing.assert_array_equal(first, second)
        else:
            assert np.abs(first[0, 0] - second[0, 0]).max() > 0
            np.testing.assert_array_equal(first[0, 1:], second[0, 1:])


def test_stft_rows_share_memory_like_one_long_stream():
    cfg = o.SpeechConfig(16000, 400, 0.01, 0.01)
    x = o.synth_signal(3, 1600, 13)  # ttd multiple of H: rows concatenate seamlessly
    two = o.stft2(x, cfg)
    cfg1 = o.SpeechConfig(16000, 400, 0.01, 0.01)
    one = o.stft2(x.reshape(1, -1), cfg1)
    # row r frame t of the 2-D call = frame r*T' + t - (stream offset) of the flat call, except the
    # flat call has no zero rows in between: compare row 1 frame 0 (needs row 0's tail)
    Tp = 10
    np.testing.assert_allclose(two[1, 0], one[0, Tp], atol=1e-7)
    with pytest.raises(ValueError):
        o.stft2(x, o.SpeechConfig(16000))  # default config: H=320 > W/2 (Q13)


def test_f32_vs_f64_oracle_within_tolerance():
    cfg = o.SpeechConfig(16000, 512, 0.025, 0.01)
    x = o.synth_signal(4, 16000, 21)
    for c in x:
        f32, e32 = o.mfe(c, cfg)
        f64, e64 = o.mfe(c, cfg, np.float64)
        assert np.abs(f32 / f64 - 1).max() < 1e-4 and np.abs(e32 / e64 - 1).max() < 1e-5
        assert np.abs(o.mfcc(c, cfg) - o.mfcc(c, cfg, np.float64)).max() < 1e-4


def test_golden_fixtures_match_oracle():
    """tests/golden/*.npz were generated by tests/golden/make_golden.py from this oracle; this pins
    the oracle against accidental drift (the reference itself holds no golden vectors)."""
    meta = json.loads((GOLD / "golden.json").read_text())
    for name, m in meta.items():
        z = np.load(GOLD / f"{name}.npz")
        cfg = o.SpeechConfig(**m["config"])
        x = o.synth_signal(m["n_clips"], m["clip_len"], m["seed"])
        if m["kind"] == "speechpy":
            np.testing.assert_allclose(np.stack([o.mfcc(c, cfg) for c in x]), z["mfcc"], atol=2e-6, rtol=1e-6)
            np.testing.assert_allclose(np.stack([o.mfe(c, cfg)[0] for c in x]), z["mfe"], rtol=2e-6)
        else:
            np.testing.assert_allclose(o.mel_spectrogram2(x, cfg), z["mel"], rtol=1e-5, atol=1e-12)


def test_next_rows_deltas_and_cmvnw():
    """closed forms for the delta features and the sliding-window CMVN"""
    f = np.arange(12, dtype=np.float32).reshape(2, 6) ** 2
    d = o.derivative_extraction(f, 2)
    F = np.pad(f, ((0, 0), (2, 2)), mode="edge")
    want = ((F[:, 3:9] * 1 - F[:, 1:7]) + (F[:, 4:10] * 2 - F[:, 0:6])) / 10.0   # Range * A - B precedence
    np.testing.assert_allclose(d, want, rtol=1e-6)
    cube = o.extract_derivative_feature(f)
    assert cube.shape == (2, 6, 3)
    np.testing.assert_array_equal(cube[:, :, 0], f)
    np.testing.assert_array_equal(cube[:, :, 2], o.derivative_extraction(cube[:, :, 1], 2))
    x = o.synth_signal(400, 5, 8) * 4 + 1
    for vn in (False, True):
        np.testing.assert_allclose(o.cmvnw(x, 301, vn), o.cmvnw_direct(x, 301, vn), atol=2e-5)
    np.testing.assert_allclose(o.cmvnw(np.full((40, 3), 2.5, np.float32), 11, False), 0, atol=1e-6)
    with pytest.raises(ValueError):
        o.cmvnw(x, 300, False)
