"""The oracle's C port (oracle/oracle_c.c, the timed CPU baseline) against the numpy restatement:
two independent implementations (own Stockham FFT vs scipy's pocketfft) must agree."""
import numpy as np
import pytest

import oracle_c
import speechsauce_oracle as o
from test_host_abi import CONFIGS


@pytest.mark.parametrize("cname", ["c1", "default", "odd"])
def test_c_port_features_match_numpy_oracle(cname):
    kw = CONFIGS[cname]
    oc, cc = o.SpeechConfig(**kw), oracle_c.make_config(**kw)
    x = o.synth_signal(6, oc.sample_rate, 17)
    x[1] = 0.0                      # zero_handling path
    x[2, :4000] = 0.0               # silence next to sound
    r = oracle_c.features(x, cc, mfe=True, energy=True, lmfe=True, mfcc=True, cmvn_mode=0, n_threads=2)
    for i in range(6):
        f, e = o.mfe(x[i], oc)
        np.testing.assert_allclose(r["mfe"][i], f, rtol=1e-4, atol=2e-6 * f.max())
        np.testing.assert_allclose(r["energy"][i], e, rtol=1e-5)
        np.testing.assert_allclose(r["lmfe"][i], np.log(f), atol=1e-3)
        np.testing.assert_allclose(r["mfcc"][i], o.mfcc(x[i], oc), atol=1e-3)
    assert (r["mfe"][1] == o.EPS).all()
    for mode in (1, 2):
        rc = oracle_c.features(x[3:], cc, mfcc=True, cmvn_mode=mode)
        for i in range(3):
            np.testing.assert_allclose(rc["mfcc"][i], o.cmvn(o.mfcc(x[3 + i], oc), mode == 2), atol=1e-3)


@pytest.mark.parametrize("cname,ttd", [("c3", 22050), ("c4", 16000), ("c4", 4001)])
def test_c_port_mel_spectrogram_matches_numpy_oracle(cname, ttd):
    kw = CONFIGS[cname]
    y = o.synth_signal(4, ttd, 23)
    a = oracle_c.mel_spectrogram2(y, oracle_c.make_config(**kw), n_threads=2)
    b = o.mel_spectrogram2(y, o.SpeechConfig(**kw))
    assert a.shape == b.shape
    np.testing.assert_allclose(a, b, rtol=1e-4, atol=2e-6 * b.max())
    al = oracle_c.mel_spectrogram2(y, oracle_c.make_config(**kw), log=True)
    np.testing.assert_allclose(al, o.log_mel_spectrogram2(y, o.SpeechConfig(**kw)), atol=1e-3)


def test_c_port_frame_counts():
    L = oracle_c.lib()
    for n, Lf, S in [(16000, 400, 160), (10**6, 320, 160), (57_600_000, 400, 160), (300, 400, 160)]:
        for zp in (0, 1):
            try:
                want = o.num_frames(n, Lf, S, bool(zp))
                want = -1 if (not zp and want < 1) else want
            except ValueError:
                want = -1
            assert L.orc_num_frames(n, Lf, S, zp) == want
