"""The host side of the drop-in boundary on a GPU: pageable caller memory through the pinned staging
rings, pinned memory DMA'd in place, frame tiles for long clips, the in-process multi-device plan, the
device guard and the (device, stream)-keyed scratch of the *_dev entry points.  What the reference's
callers hand over is pageable ndarray memory (py-speechsauce/src/lib.rs:167-204)."""
import ctypes as C

import numpy as np
import pytest

import speechsauce_b200 as sb
import speechsauce_oracle as o
from conftest import assert_log_close
from speechsauce_b200 import _lib, hostmem
from test_host_abi import CONFIGS

pytestmark = pytest.mark.gpu

ALL = dict(mfe=True, energy=True, lmfe=True, ssc=True, mfcc=True)


def _pinned_copy(x):
    p = hostmem.pinned_empty(x.shape, x.dtype)
    p[...] = x
    return p


def test_pinned_and_pageable_host_buffers_give_identical_results():
    x = o.synth_signal(700, 16000, 501)
    cfg = sb.SpeechConfig(**CONFIGS["c1"], devices=[0])
    assert not hostmem.is_pinned(x)
    a = sb.feature.features(x, cfg, cmvn_mode=2, **ALL)                      # pageable in, pageable out
    xp = _pinned_copy(x)
    assert hostmem.is_pinned(xp)
    b = sb.feature.features(xp, cfg, cmvn_mode=2, empty=hostmem.pinned_empty, **ALL)  # pinned both ways
    c = sb.feature.features(xp, cfg, cmvn_mode=2, **ALL)                     # pinned in, pageable out
    for k in a:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
        np.testing.assert_array_equal(a[k], c[k], err_msg=k)
    # and it is the right answer (first and last clip against the oracle)
    oc = o.SpeechConfig(**CONFIGS["c1"])
    for i in (0, 699):
        assert_log_close(a["mfcc"][i], o.cmvn(o.mfcc(x[i], oc), True), what=f"clip {i}")
    # registering a caller-owned array pins it in place
    y = x.copy()
    hostmem.pin(y)
    try:
        assert hostmem.is_pinned(y)
        d = sb.feature.features(y, cfg, cmvn_mode=2, **ALL)
    finally:
        hostmem.unpin(y)
    for k in a:
        np.testing.assert_array_equal(a[k], d[k], err_msg=k)


@pytest.mark.parametrize("chunk_bytes", [200_000, 3_000_000])
def test_small_chunks_reuse_the_staging_rings_correctly(chunk_bytes):
    """many chunks per lane: a ring slot is refilled only after its previous rows reached the caller"""
    x = o.synth_signal(257, 16000, 502)
    ref = sb.feature.features(x, sb.SpeechConfig(**CONFIGS["c1"], devices=[0]), cmvn_mode=1, **ALL)
    cfg = sb.SpeechConfig(**CONFIGS["c1"], devices=[0]).set_tuning(chunk_bytes=chunk_bytes)
    got = sb.feature.features(x, cfg, cmvn_mode=1, **ALL)
    for k in ref:
        np.testing.assert_array_equal(ref[k], got[k], err_msg=k)  # the cmvn kernel is chosen by the clip shape, not the chunking
    # strided rows (row_stride > clip_len), pageable and pinned
    wide = np.zeros((257, 16000 + 37), np.float32)
    wide[:, :16000] = x
    for buf in (wide, _pinned_copy(wide)):
        out = np.empty_like(ref["mfcc"])
        ob = _lib.ss_feature_outputs(None, None, None, None, _lib.fptr(out), 1, 0)
        _lib.check(sb.lib.ss_features(cfg.plan, _lib.fptr(buf), 257, 16000, 16037, C.byref(ob)))
        np.testing.assert_array_equal(ref["mfcc"], out)
    # STFT family through the same rings (the halo of every chunk is restated on the host)
    y = o.synth_signal(40, 8000, 503)
    c4 = sb.SpeechConfig(**CONFIGS["c4"], devices=[0])
    want = sb.functions.stft2(y, c4)
    c4s = sb.SpeechConfig(**CONFIGS["c4"], devices=[0]).set_tuning(chunk_bytes=chunk_bytes)
    np.testing.assert_array_equal(want, sb.functions.stft2(y, c4s))
    np.testing.assert_array_equal(sb.feature.mel_spectrogram2(y, c4), sb.feature.mel_spectrogram2(_pinned_copy(y), c4s))


@pytest.mark.parametrize("cmvn_mode", [0, 1, 2])
def test_frame_tiles_of_a_long_clip_match_the_whole_clip(cmvn_mode):
    """one long stream is cut into frame tiles with an L - S halo; the length-dependent DCT scale
    (feature.rs:126), the [0,0]-only k = 0 scale (feature.rs:128) and the clip-wide cmvn must survive"""
    x = o.synth_signal(3, 16000 * 21 + 123, 504)
    whole = sb.SpeechConfig(**CONFIGS["c1"], devices=[0])
    ref = sb.feature.features(x, whole, cmvn_mode=cmvn_mode, **ALL)
    T = ref["mfcc"].shape[1]
    for tile in (100, 777, T - 1, T):          # ragged last tile, one-frame last tile, exactly one tile
        cfg = sb.SpeechConfig(**CONFIGS["c1"], devices=[0]).set_tuning(tile_frames=tile)
        got = sb.feature.features(x, cfg, cmvn_mode=cmvn_mode, **ALL)
        for k in ("mfe", "energy", "lmfe", "ssc"):
            np.testing.assert_array_equal(ref[k], got[k], err_msg=f"{k} tile={tile}")
        if cmvn_mode == 0:
            np.testing.assert_array_equal(ref["mfcc"], got["mfcc"], err_msg=f"tile={tile}")
        else:
            np.testing.assert_allclose(got["mfcc"], ref["mfcc"], rtol=0, atol=2e-5, err_msg=f"tile={tile}")
    # pinned input and outputs: the tiles (overlapping rows) are DMA'd in place
    cfg = sb.SpeechConfig(**CONFIGS["c1"], devices=[0]).set_tuning(tile_frames=300)
    got = sb.feature.features(_pinned_copy(x), cfg, cmvn_mode=cmvn_mode, empty=hostmem.pinned_empty, **ALL)
    for k in ("mfe", "energy", "lmfe", "ssc"):
        np.testing.assert_array_equal(ref[k], got[k], err_msg=f"{k} pinned tiles")
    np.testing.assert_allclose(got["mfcc"], ref["mfcc"], rtol=0, atol=2e-5 if cmvn_mode else 0)
    # dc_elimination off: column 0 keeps the reference's frame-0-only scale across tiles
    kw = dict(CONFIGS["c1"], dc_elimination=False)
    ref = sb.feature.features(x, sb.SpeechConfig(**kw, devices=[0]), mfcc=True)["mfcc"]
    got = sb.feature.features(x, sb.SpeechConfig(**kw, devices=[0]).set_tuning(tile_frames=500), mfcc=True)["mfcc"]
    np.testing.assert_array_equal(ref, got)
    oc = o.SpeechConfig(**kw)
    assert_log_close(got[1], o.mfcc(x[1], oc), atol=4e-3, what="tiled mfcc, dc_elimination off")


def test_multi_device_plan_pageable_pinned_and_tiled():
    n = sb.lib.ss_device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    devs = list(range(n))
    x = o.synth_signal(1001, 16000, 505)
    one = sb.feature.features(x, sb.SpeechConfig(**CONFIGS["c1"], devices=[0]), cmvn_mode=2, **ALL)
    cfg = sb.SpeechConfig(**CONFIGS["c1"], devices=devs)
    for rep in range(3):  # the worker threads persist across calls
        many = sb.feature.features(x, cfg, cmvn_mode=2, **ALL)
        for k in one:
            np.testing.assert_array_equal(one[k], many[k], err_msg=f"{k} rep {rep}")
    many = sb.feature.features(_pinned_copy(x), cfg, cmvn_mode=2, empty=hostmem.pinned_empty, **ALL)
    for k in one:
        np.testing.assert_array_equal(one[k], many[k], err_msg=k)
    # fewer clips than devices, and long clips in tiles on several devices
    z = o.synth_signal(n + 1, 16000 * 9, 506)
    a = sb.feature.features(z, sb.SpeechConfig(**CONFIGS["c1"], devices=[0]), cmvn_mode=2, **ALL)
    b = sb.feature.features(z, sb.SpeechConfig(**CONFIGS["c1"], devices=devs).set_tuning(tile_frames=250, chunk_bytes=1 << 20),
                            cmvn_mode=2, **ALL)
    for k in ("mfe", "energy", "lmfe", "ssc"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    np.testing.assert_allclose(b["mfcc"], a["mfcc"], rtol=0, atol=2e-5)
    one_clip = sb.feature.features(z[:1], sb.SpeechConfig(**CONFIGS["c1"], devices=devs), mfcc=True)["mfcc"]
    np.testing.assert_array_equal(one_clip, sb.feature.features(z[:1], sb.SpeechConfig(**CONFIGS["c1"], devices=[0]), mfcc=True)["mfcc"])
    # STFT rows over several devices: the row-to-row overlap memory (Q13) crosses the device boundary as a host halo
    y = o.synth_signal(4 * n + 3, 8000, 507)
    c41 = sb.SpeechConfig(**CONFIGS["c4"], devices=[0])
    c4n = sb.SpeechConfig(**CONFIGS["c4"], devices=devs).set_tuning(chunk_bytes=1 << 16)
    np.testing.assert_array_equal(sb.functions.stft2(y, c41), sb.functions.stft2(y, c4n))
    np.testing.assert_array_equal(sb.feature.mel_spectrogram2(y, c41), sb.feature.mel_spectrogram2(y, c4n))
    np.testing.assert_array_equal(c41.analysis_mem, c4n.analysis_mem)


def test_host_calls_leave_the_callers_current_device_alone():
    torch = pytest.importorskip("torch")
    n = sb.lib.ss_device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    torch.cuda.set_device(0)
    torch.zeros(1, device="cuda")
    x = o.synth_signal(5, 16000, 508)
    cfg = sb.SpeechConfig(**CONFIGS["c1"], devices=[1])
    sb.feature.features(x, cfg, mfcc=True)
    assert torch.cuda.current_device() == 0
    t = torch.ones(4, device="cuda")
    assert t.device.index == 0
    y = o.synth_signal(2, 8000, 509)
    sb.functions.stft2(y, sb.SpeechConfig(**CONFIGS["c4"], devices=[1]))
    assert torch.cuda.current_device() == 0


def test_dev_entry_points_on_two_streams_do_not_share_scratch():
    """ss_cmvn_dev / ss_cmvnw_dev kept one scratch buffer per thread, so two streams raced"""
    torch = pytest.importorskip("torch")
    g = torch.Generator(device="cpu").manual_seed(5)
    mats = [torch.randn(64, 4000, 13, generator=g).cuda() for _ in range(2)]
    outs = [torch.empty_like(m) for m in mats]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    for rep in range(4):
        for m, out, s in zip(mats, outs, streams):
            _lib.check(sb.lib.ss_cmvn_dev(C.c_void_p(s.cuda_stream), C.c_void_p(m.data_ptr()), 64, 4000, 13, 1,
                                          C.c_void_p(out.data_ptr())))
            _lib.check(sb.lib.ss_cmvnw_dev(C.c_void_p(s.cuda_stream), C.c_void_p(m.data_ptr()), 64, 4000, 13, 301, 1,
                                           C.c_void_p(out.data_ptr())))
            _lib.check(sb.lib.ss_cmvn_dev(C.c_void_p(s.cuda_stream), C.c_void_p(m.data_ptr()), 64, 4000, 13, 1,
                                          C.c_void_p(out.data_ptr())))
    torch.cuda.synchronize()
    for m, out in zip(mats, outs):
        want = o.cmvn(m[3].cpu().numpy(), True)
        np.testing.assert_allclose(out[3].cpu().numpy(), want, rtol=0, atol=2e-5)
    # power_spectrum's twiddle table is uploaded on the first caller's stream; a second stream must wait for it
    fr = torch.randn(2000, 300).cuda()
    ps = [torch.empty(2000, 301, device="cuda") for _ in range(2)]
    torch.cuda.synchronize()
    for p_, s in zip(ps, streams):
        _lib.check(sb.lib.ss_power_spectrum_dev(C.c_void_p(s.cuda_stream), C.c_void_p(fr.data_ptr()), 2000, 300, 600,
                                                C.c_void_p(p_.data_ptr())))
    torch.cuda.synchronize()
    assert torch.equal(ps[0], ps[1])
    np.testing.assert_allclose(ps[0][5].cpu().numpy(), o.power_spectrum(fr[5:6].cpu().numpy(), 600)[0], rtol=1e-4, atol=1e-6)
    assert sb.lib.ss_release_scratch() == 0


def test_device_arrays_in_and_out_without_a_host_round_trip():
    """`__cuda_array_interface__` inputs and outputs (torch tensors here, and a bare object
    that carries nothing but the interface dict, standing in for cupy / numba)"""
    torch = pytest.importorskip("torch")
    x = o.synth_signal(9, 16000, 510)
    xd = torch.from_numpy(x).cuda()
    cfg = sb.SpeechConfig(**CONFIGS["c1"], devices=[0])
    host = sb.feature.features(x, cfg, cmvn_mode=2, **ALL)
    dev = sb.device.features(xd, cfg, cmvn_mode=2, **ALL)
    torch.cuda.synchronize()
    for k in host:
        assert dev[k].is_cuda and dev[k].__cuda_array_interface__["data"][0] == dev[k].data_ptr()
        np.testing.assert_array_equal(host[k], dev[k].cpu().numpy(), err_msg=k)

    class Bare:  # only the protocol
        def __init__(self, t):
            self._t = t
            self.__cuda_array_interface__ = dict(t.__cuda_array_interface__)

    s = torch.cuda.Stream()
    one = sb.device.mfcc(Bare(xd[3]), cfg, stream=s)
    s.synchronize()
    np.testing.assert_array_equal(one.cpu().numpy(), sb.feature.mfcc(x[3], cfg))
    y = o.synth_signal(3, 8000, 511)
    c4 = sb.SpeechConfig(**CONFIGS["c4"], devices=[0])
    m = sb.device.mel_spectrogram2(torch.from_numpy(y).cuda(), c4, log=True)
    st = sb.device.stft2(torch.from_numpy(y).cuda(), c4)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(m.cpu().numpy(), sb.feature.mel_spectrogram2(y, sb.SpeechConfig(**CONFIGS["c4"], devices=[0]), log=True))
    np.testing.assert_array_equal(torch.view_as_complex(st).cpu().numpy(), sb.functions.stft2(y, sb.SpeechConfig(**CONFIGS["c4"], devices=[0])))
    f = torch.randn(5, 97, 13, device="cuda")
    np.testing.assert_allclose(sb.device.cmvn(f, True).cpu().numpy()[2], o.cmvn(f[2].cpu().numpy(), True), atol=2e-5)
    p = sb.device.preemphasis(xd[0])
    np.testing.assert_array_equal(p.cpu().numpy(), o.preemphasis(x[0], 1, 0.98))
    with pytest.raises(ValueError):
        sb.device.features(xd.t(), cfg, mfcc=True)          # not C-contiguous
    with pytest.raises(TypeError):
        sb.device.features(xd.double(), cfg, mfcc=True)     # f32 only
    with pytest.raises(TypeError):
        sb.device.features(x, cfg, mfcc=True)               # a host array


def test_analysis_mem_arguments_are_validated():
    y = o.synth_signal(2, 8000, 512)
    c4 = sb.SpeechConfig(**CONFIGS["c4"], devices=[0])
    good = np.zeros(240, np.float32)
    sb.functions.stft2(y, c4, analysis_mem=good)
    for bad in (np.zeros(240, np.float64), np.zeros(239, np.float32), np.zeros(480, np.float32)[::2], [0.0] * 240):
        with pytest.raises(ValueError):
            sb.functions.stft2(y, c4, analysis_mem=bad)
        with pytest.raises(ValueError):
            sb.feature.mel_spectrogram2(y, c4, analysis_mem=bad)
