"""Independent pins of the oracle's sub-steps: code the builder did not write (torch.stft / torch.fft on MKL,
torchaudio, scipy) against oracle/speechsauce_oracle.py.  The reference holds no golden vectors
(speechsauce/src/lib.rs:41-134 asserts shapes only) and cannot be built here, so these are the strongest pins
available for the pieces whose definition is standard: the real FFT magnitude, the windowed STFT, the DCT-II.
What stays pinned by the restatement alone: the reference-specific composition (frame counts without "+1", the
half-coverage filterbank, the T*nf DCT scale, the STFT's overlap-memory stream model)."""
import numpy as np
import pytest
import scipy.fft

import speechsauce_oracle as o
from test_host_abi import CONFIGS

torch = pytest.importorskip("torch")


def test_power_spectrum_against_torch_rfft_and_torchaudio():
    """processing::power_spectrum = |rfft(frame, n_fft)| / n_fft (processing.rs:143-181)"""
    rng = np.random.default_rng(1)
    for L, n_fft in [(400, 512), (320, 512), (512, 512), (400, 400), (2048, 2048), (77, 128)]:
        fr = (rng.standard_normal((9, L)) * 0.1).astype(np.float32)
        got = o.power_spectrum(fr, n_fft)
        want = torch.fft.rfft(torch.from_numpy(fr).double(), n=n_fft, dim=1).abs().numpy() / n_fft
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=2e-6 * want.max())
    torchaudio = pytest.importorskip("torchaudio")
    # torchaudio's Spectrogram(power=1) over one un-centred rectangular frame is the same |rfft|
    fr = (rng.standard_normal((5, 512)) * 0.1).astype(np.float32)
    spec = torchaudio.transforms.Spectrogram(n_fft=512, win_length=512, hop_length=512, window_fn=torch.ones, power=1.0, center=False)
    want = spec(torch.from_numpy(fr))[:, :, 0].numpy() / 512
    np.testing.assert_allclose(o.power_spectrum(fr, 512), want, rtol=1e-4, atol=2e-6 * want.max())


@pytest.mark.parametrize("cname,lead", [("c3", 0), ("c4", 80)])
def test_stft_against_torch_stft_with_the_vorbis_window(cname, lead):
    """functions::stft2 (functions.rs:86-170) on one fresh row: output row t is the windowed DFT of the W samples that
    end at (t + n_pad + 1) H.  With H | W (c3) that is torch.stft(center=False) frame t; with W = 400, H = 160 (c4) the
    frame starts W - (n_pad + 1) H = 80 samples early, inside the (zero) overlap memory."""
    cfg = o.SpeechConfig(**CONFIGS[cname])
    W, H, n_pad = cfg.window_size, cfg.frame_size, cfg.window_size // cfg.frame_size - 1
    assert W - (n_pad + 1) * H == lead
    x = o.synth_signal(1, 7 * W + 13, 5)[0]
    got = o.stft2(x[None, :], cfg)[0]                       # [T', F] complex
    Tp = got.shape[0]
    padded = np.concatenate([np.zeros(lead, np.float32), x, np.zeros(Tp * H - len(x), np.float32)])
    win = torch.from_numpy(np.asarray(cfg.window, np.float64))
    ref = torch.stft(torch.from_numpy(padded).double(), n_fft=W, hop_length=H, win_length=W, window=win, center=False,
                     return_complex=True).numpy().T * float(cfg.wnorm)
    n = Tp - n_pad
    assert ref.shape[0] >= n
    scale = np.abs(ref[:n]).max()
    assert np.abs(got[:n] - ref[:n]).max() <= 2e-6 * scale  # f32 FFT noise floor relative to the peak
    np.testing.assert_allclose(np.abs(got[:n]), np.abs(ref[:n]), rtol=1e-4, atol=2e-6 * scale)
    assert (got[n:] == 0).all()                             # the reference's trailing zero rows (functions.rs:121)


def test_vorbis_window_closed_form_and_power_complementarity():
    """config.rs:151-160: w[i] = sin(pi/2 sin^2(pi (i + 1/2) / N)); a Vorbis window satisfies w[i]^2 + w[i + N/2]^2 = 1"""
    for cname in ("c3", "c4"):
        cfg = o.SpeechConfig(**CONFIGS[cname])
        w = np.asarray(cfg.window, np.float64)
        N = len(w)
        np.testing.assert_allclose(w[: N // 2] ** 2 + w[N // 2:] ** 2, 1.0, atol=1e-6)
        i = np.arange(N)
        np.testing.assert_allclose(w, np.sin(0.5 * np.pi * np.sin(np.pi * (i + 0.5) / N) ** 2), atol=1e-7)


def test_dct2_against_scipy_and_torch():
    """nddct2 (feature.rs:120-123) is assumed to be scipy's unnormalised DCT-II, 2 sum x[n] cos(pi k (2n + 1) / 2N)"""
    rng = np.random.default_rng(2)
    x = rng.standard_normal((11, 40))
    k, n = np.meshgrid(np.arange(40), np.arange(40), indexing="ij")
    direct = 2.0 * (x[:, None, :] * np.cos(np.pi * k * (2 * n + 1) / 80.0)[None]).sum(-1)
    np.testing.assert_allclose(scipy.fft.dct(x, type=2, axis=1), direct, atol=1e-10)
    # the same numbers from an FFT-based construction in torch (real part of a twiddled FFT of the mirrored sequence)
    xt = torch.from_numpy(x)
    y = torch.fft.fft(torch.cat([xt, xt.flip(1)], dim=1), dim=1)[:, :40]
    tw = torch.exp(-1j * np.pi * torch.arange(40, dtype=torch.float64) / 80.0)
    np.testing.assert_allclose((y * tw).real.numpy(), direct, atol=1e-9)
    # and the oracle's mfcc uses exactly this transform on ln(mfe): cepstra 1.. of one clip, Q9 scale undone
    cfg = o.SpeechConfig(**CONFIGS["c1"])
    sig = o.synth_signal(1, 16000, 9)[0]
    feat, _ = o.mfe(sig, cfg)
    T, nf = feat.shape
    cep = scipy.fft.dct(np.log(feat.astype(np.float64)), type=2, axis=1)[:, :13] / np.sqrt(2.0 * T * nf)
    np.testing.assert_allclose(o.mfcc(sig, cfg)[:, 1:], cep[:, 1:], atol=1e-3)
