"""GPU tests at BASELINE.json's full sizes.  The oracle cannot process whole batches in reasonable
time, so besides oracle spot checks on sampled clips these tests use size-independent properties of
the path: scale equivariance (a power-of-two gain is exact in every FFT butterfly), batch-position
independence, time-shift equivariance of the framing, and the CMVN column statistics reduced over
the whole batch on the device.  Device-resident through the *_dev entry points (no 4-46 GB host
copies); inputs are generated on the device like bench.py does."""
import ctypes as C
import math

import numpy as np
import pytest

import speechsauce_b200 as sb
import speechsauce_oracle as o
from conftest import assert_log_close, assert_spec_close
from speechsauce_b200 import _lib
from test_host_abi import CONFIGS

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _features_dev(cfg, x, outs, cmvn):
    n, length = x.shape
    o_ = _lib.ss_feature_outputs(*[C.c_void_p(outs[k].data_ptr()) if k in outs else None
                                   for k in ("mfe", "energy", "lmfe", "ssc", "mfcc")], cmvn, 0)
    _lib.check(sb.lib.ss_features_dev(cfg.plan, 0, _stream(), C.c_void_p(x.data_ptr()), n, length, x.stride(0), C.byref(o_)))
    torch.cuda.synchronize()


def _mel_dev(cfg, x, out, log):
    n, ttd = x.shape
    _lib.check(sb.lib.ss_mel_spectrogram_dev(cfg.plan, 0, _stream(), C.c_void_p(x.data_ptr()), n, ttd, x.stride(0),
                                             C.c_void_p(out.data_ptr()), None, None, log))
    torch.cuda.synchronize()


def _need(gb):
    free, _ = torch.cuda.mem_get_info()
    if free < gb * (1 << 30):
        pytest.skip(f"needs {gb} GB of free device memory")


def test_c2_full_batch_properties():
    """65,536 x 1 s clips -> lmfe + cmvn(mfcc): gain equivariance, position independence, CMVN statistics."""
    _need(16)
    n, length, T = 65536, 16000, 97
    cfg = sb.SpeechConfig(**CONFIGS["c1"], devices=[0])
    oc = o.SpeechConfig(**CONFIGS["c1"])
    gen = torch.Generator(device="cuda")
    gen.manual_seed(0x5EED0002)
    x = torch.empty((n, length), dtype=torch.float32, device="cuda").normal_(0.0, 0.1, generator=gen)
    a = {"lmfe": torch.empty((n, T, 40), dtype=torch.float32, device="cuda"),
         "mfcc": torch.empty((n, T, 13), dtype=torch.float32, device="cuda")}
    _features_dev(cfg, x, a, 2)
    assert sb.lib.ss_last_kernel_name() == b"k_fast512c"
    assert torch.isfinite(a["lmfe"]).all() and torch.isfinite(a["mfcc"]).all()
    # CMVN over every clip of the batch (processing.rs:265-300): column mean 0, std 1
    assert a["mfcc"].mean(1).abs().max().item() < 2e-5
    assert (a["mfcc"].std(1, unbiased=False) - 1).abs().max().item() < 1e-4
    # a gain of 4 scales every mel energy by exactly 4: ln shifts by ln 4, cmvn(mfcc) does not move
    b = {k: torch.empty_like(v) for k, v in a.items()}
    _features_dev(cfg, x * 4.0, b, 2)
    assert (b["lmfe"] - a["lmfe"] - math.log(4.0)).abs().max().item() < 1e-5
    assert (b["mfcc"] - a["mfcc"]).abs().max().item() < 1e-3
    # clips computed alone (another launch shape: chunked items + the standalone CMVN kernels) match their rows
    pick = torch.tensor([0, 1, 777, 12345, 40000, 65535], device="cuda")
    small = {"lmfe": torch.empty((len(pick), T, 40), dtype=torch.float32, device="cuda"),
             "mfcc": torch.empty((len(pick), T, 13), dtype=torch.float32, device="cuda")}
    _features_dev(cfg, x[pick].contiguous(), small, 2)
    assert torch.equal(small["lmfe"], a["lmfe"][pick])
    assert (small["mfcc"] - a["mfcc"][pick]).abs().max().item() < 1e-4
    # and the oracle on those clips
    xs = x[pick].cpu().numpy()
    for j in range(len(pick)):
        assert_log_close(small["lmfe"][j].cpu().numpy(), o.lmfe(xs[j], oc), what=f"c2 clip {int(pick[j])} lmfe")
        assert_log_close(a["mfcc"][pick[j]].cpu().numpy(), o.cmvn(o.mfcc(xs[j], oc), True), what=f"c2 clip {int(pick[j])} cmvn(mfcc)")


def test_c5_hour_long_streams_properties():
    """128 one-hour streams -> mfe + energy + ssc + mfcc: a stream cut at a frame boundary gives the same frames."""
    _need(60)
    n, length = 128, 57_600_000
    T = o.num_frames(length, 400, 160, False)
    assert T == 359_997
    cfg = sb.SpeechConfig(**CONFIGS["c1"], devices=[0])
    oc = o.SpeechConfig(**CONFIGS["c1"])
    gen = torch.Generator(device="cuda")
    gen.manual_seed(0x5EED0005)
    x = torch.empty((n, length), dtype=torch.float32, device="cuda")
    for i in range(0, n, 8):
        x[i:i + 8].normal_(0.0, 0.1, generator=gen)
    dims = {"mfe": (40,), "energy": (), "ssc": (40,), "mfcc": (13,)}
    a = {k: torch.empty((n, T) + d, dtype=torch.float32, device="cuda") for k, d in dims.items()}
    _features_dev(cfg, x, a, 0)
    assert all(torch.isfinite(v).all() for v in a.values())
    # framing is memoryless: the tail of a stream from frame f0 on, computed alone, reproduces frames f0.. —
    # except that mfcc's 1/sqrt(2 T nf) scale depends on the number of frames (feature.rs:126, quirk Q9)
    f0, nt = 200_003, 5000
    seg = x[:4, f0 * 160: f0 * 160 + (nt - 1) * 160 + 400 + 160].contiguous()   # nt frames (no "+1" in the count, Q2)
    assert o.num_frames(seg.shape[1], 400, 160, False) == nt
    b = {k: torch.empty((4, nt) + d, dtype=torch.float32, device="cuda") for k, d in dims.items()}
    _features_dev(cfg, seg, b, 0)
    for k in ("mfe", "energy", "ssc"):
        assert torch.equal(b[k], a[k][:4, f0:f0 + nt]), k
    assert torch.equal(b["mfcc"][..., 0], a["mfcc"][:4, f0:f0 + nt, 0])        # c0 = ln(energy): no scale
    ratio = math.sqrt(T / nt)
    assert (b["mfcc"][..., 1:] - a["mfcc"][:4, f0:f0 + nt, 1:] * ratio).abs().max().item() < 1e-5
    # oracle on a 20 s excerpt of one stream
    ex = x[77, :320_000].cpu().numpy()
    Te = o.num_frames(ex.size, 400, 160, False)
    f, e = o.mfe(ex, oc)
    assert_spec_close(a["mfe"][77, :Te].cpu().numpy(), f, what="c5 mfe excerpt")
    assert_spec_close(a["energy"][77, :Te].cpu().numpy(), e, what="c5 energy excerpt")
    assert_spec_close(a["ssc"][77, :Te].cpu().numpy(), o.ssc(ex, oc), what="c5 ssc excerpt")
    got, want = a["mfcc"][77, :Te].cpu().numpy(), o.mfcc(ex, oc)
    assert_log_close(got[:, 0], want[:, 0], what="c5 mfcc c0 excerpt")
    assert_log_close(got[:, 1:] * math.sqrt(T / Te), want[:, 1:], what="c5 mfcc excerpt (rescaled to the excerpt's frame count)")


@pytest.mark.parametrize("cname,n,ttd,log,gb", [("c4", 8192, 480000, 1, 40), ("c3", 4096, 110250, 0, 6)])
def test_stft_full_batch_properties(cname, n, ttd, log, gb):
    """c4 (8,192 x 30 s, log-mel) and c3 (4,096 x 5 s): dropping the first hop of every row shifts the interior
    frames by one; rows computed alone match their place in the batch; oracle on sampled rows."""
    _need(gb)
    kw = CONFIGS[cname]
    cfg = sb.SpeechConfig(**kw, devices=[0])
    oc = o.SpeechConfig(**kw)
    H, nf = oc.frame_size, oc.num_filters
    Tp = o.stft_num_frames(ttd, H)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(0x5EED0000 + int(cname[1]))
    x = torch.empty((n, ttd), dtype=torch.float32, device="cuda")
    step = max(1, (1 << 28) // ttd)
    for i in range(0, n, step):
        x[i:i + step].normal_(0.0, 0.1, generator=gen)
    a = torch.empty((n, nf, Tp), dtype=torch.float32, device="cuda")
    _mel_dev(cfg, x, a, log)
    assert sb.lib.ss_last_kernel_name() in (b"k_stft400m", b"k_stft2048m")
    assert torch.isfinite(a).all()
    n_pad = oc.window_size // H - 1
    tail = a[:, :, Tp - n_pad:]                 # the reference's zero rows (Q13), ln(EPS) in log mode
    assert (tail - (math.log(o.EPS) if log else 0.0)).abs().max().item() <= (1e-5 if log else 0.0)
    # rows 64.. shifted by one hop (a strided view of the same buffer): interior frames move by one
    m = 64
    xs = x[:m, H:]
    b = torch.empty((m, nf, o.stft_num_frames(ttd - H, H)), dtype=torch.float32, device="cuda")
    _mel_dev(cfg, xs, b, log)
    lo, hi = n_pad + 2, b.shape[2] - n_pad - 2   # frames that touch neither the overlap memory nor the padding
    assert torch.equal(b[:, :, lo:hi], a[:m, :, lo + 1:hi + 1])
    # sampled rows computed alone (row 0 of a call has zero overlap memory: compare from the second frame on
    # unless the row is the batch's first) and the oracle on them
    for r in (0, n // 2 + 1, n - 1):
        one = torch.empty((1, nf, Tp), dtype=torch.float32, device="cuda")
        _mel_dev(cfg, x[r:r + 1], one, log)
        assert torch.equal(one[0, :, 1:], a[r, :, 1:])
        want = o.mel_spectrogram2(x[r:r + 1].cpu().numpy(), o.SpeechConfig(**kw))[0]
        got = one[0].cpu().numpy()
        if log:
            # ln(zero_handling(mel)): exact zeros (empty filters, zero tail) -> ln EPS; abs 1e-3 on everything above
            # the f32 noise floor of the transform.  Over 240,000 values per row a few single-bin filters land 9
            # orders of magnitude below the frame's peak, where the f32 oracle and the f64 oracle differ by 1e-3
            # themselves; those are checked as spectra (relative + absolute floor) instead.
            zero = want == 0
            assert np.abs(got[zero] - math.log(o.EPS)).max(initial=0.0) < 1e-5
            loud = want > 1e-6 * want.max()
            assert loud.mean() > 0.7
            assert_log_close(got[loud], np.log(want[loud]), what=f"{cname} row {r} log-mel")
            assert_spec_close(np.where(zero, 0.0, np.exp(got.astype(np.float64))), want, what=f"{cname} row {r} exp(log-mel)")
        else:
            assert_spec_close(got, want, what=f"{cname} row {r} mel")
