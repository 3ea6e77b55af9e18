"""Small invocation of every kernel for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool memcheck python tests/sanitize_small.py
Not a pytest file; shapes are tiny so the instrumented run finishes in seconds."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "mfcc-rust_b200"), str(ROOT / "oracle")]
import numpy as np  # noqa: E402
import speechsauce_b200 as sb  # noqa: E402
import speechsauce_oracle as o  # noqa: E402
from speechsauce_b200 import _lib  # noqa: E402

C1 = dict(sample_rate=16000, fft_points=512, frame_length=0.025, frame_stride=0.01, num_cepstral=13, num_filters=40,
          low_frequency=0.0, high_frequency=8000.0)
C3 = dict(sample_rate=22050, fft_points=2048, frame_length=512 / 22050, frame_stride=0.01, num_cepstral=13,
          num_filters=128, low_frequency=0.0, high_frequency=11025.0)
C4 = dict(sample_rate=16000, fft_points=400, frame_length=0.01, frame_stride=0.01, num_cepstral=13, num_filters=80,
          low_frequency=0.0, high_frequency=8000.0)
x = o.synth_signal(5, 16000, 1)
for flags in (0, _lib.SS_FLAG_FORCE_GENERIC):
    cfg = sb.SpeechConfig(**C1, devices=[0], flags=flags)
    sb.feature.features(x, cfg, mfe=True, energy=True, lmfe=True, ssc=True, mfcc=True, cmvn_mode=2)
    sb.feature.features(x[:, :5003], cfg, lmfe=True, mfcc=True, cmvn_mode=0)
    sb.feature.features(o.synth_signal(2, 200_000, 2), cfg, mfcc=True, cmvn_mode=2)  # too long for the fused stash
    for c, n, ttd in ((C4, 3, 4001), (C4, 2, 100), (C3, 2, 9000), (C3, 2, 700)):
        k = sb.SpeechConfig(**c, devices=[0], flags=flags)
        y = o.synth_signal(n, ttd, 3)
        sb.feature.mel_spectrogram2(y, k, log=True)
        sb.functions.stft2(y, k)
# the fused in-kernel CMVN of k_fast512c needs one clip per resident warp: device-resident batch
import ctypes as C  # noqa: E402

import torch  # noqa: E402
cfg = sb.SpeechConfig(**C1, devices=[0])
nb = 3100
dx = torch.from_numpy(o.synth_signal(nb, 16000, 7)).cuda()
lm = torch.empty((nb, 97, 40), dtype=torch.float32, device="cuda")
mf = torch.empty((nb, 97, 13), dtype=torch.float32, device="cuda")
outs = _lib.ss_feature_outputs(None, None, C.c_void_p(lm.data_ptr()), None, C.c_void_p(mf.data_ptr()), 2, 0)
_lib.check(sb.lib.ss_features_dev(cfg.plan, 0, C.c_void_p(torch.cuda.current_stream().cuda_stream), C.c_void_p(dx.data_ptr()),
                                  nb, 16000, 16000, C.byref(outs)))
torch.cuda.synchronize()
# other filter / cepstrum counts (fourth m-tile, odd nf, 16 cepstra) and a longer stream (chunked items)
for nf, nc in ((64, 16), (41, 8), (17, 12)):
    k = sb.SpeechConfig(**dict(C1, num_filters=nf, num_cepstral=nc), devices=[0])
    sb.feature.features(x[:2], k, mfe=True, energy=True, lmfe=True, ssc=True, mfcc=True)
sb.feature.features(o.synth_signal(1, 16000 * 20 + 3, 8), cfg, mfe=True, ssc=True, mfcc=True)
k4 = sb.SpeechConfig(**C4, devices=[0])
sb.feature.mel_spectrogram2(o.synth_signal(3, 16000 * 3, 9), k4, log=True)   # staged rounds + ring wrap-around
# k_stft2048m with more items than resident groups: the TMA-staged frames of a group's next item, odd rows 8-byte aligned
sb.feature.mel_spectrogram2(o.synth_signal(60, 22050, 12), sb.SpeechConfig(**C3, devices=[0]), log=True)
for hop in (160, 100):  # librosa mode on the register-FFT kernel (full-spectrum tile; staged and per-sample rounds)
    sb.feature.melspectrogram_librosa(o.synth_signal(3, 16000 * 2 + 5, 10), sb.feature.librosa_config(16000, 400, hop, 80, devices=[0]), db=True)
sb.feature.melspectrogram_librosa(o.synth_signal(2, 5000, 11), sb.feature.librosa_config(22050, 512, 128, 40, devices=[0]))
sb.processing.log_power_spectrum(o.synth_signal(9, 400, 4), 512, True)
sb.processing.preemphasis(x[0]); sb.processing.preemphasis(x)
sb.processing.stack_frames(x[0], 16000, 0.025, 0.01, None, True)
sb.processing.power_spectrum(o.synth_signal(9, 400, 4), 512)
sb.processing.cmvn(o.synth_signal(300, 13, 5), True)
sb.processing.cmvn(o.synth_signal(4, 97 * 13, 6).reshape(4, 97, 13), True)
# round 2: k_pspec512 (every pruning class), 16-byte stack_frames, one-warp cmvn, long cmvn with the merge kernel,
# pinned buffers, frame tiles, the device-array API
for L in (1, 33, 320, 400, 512):
    sb.processing.power_spectrum(o.synth_signal(5, L, 20 + L), 512)
sb.processing.stack_frames(o.synth_signal(1, 16000, 30)[0], 16000, 0.025, 0.01, None, False)
sb.processing.stack_frames(o.synth_signal(1, 1003, 31)[0], 16000, 0.025, 0.01, None, True)
sb.processing.cmvn(o.synth_signal(70, 97 * 13, 32).reshape(70, 97, 13), True)
sb.processing.cmvn(o.synth_signal(3, 50 * 100, 33).reshape(3, 50, 100), False)
sb.processing.cmvn(o.synth_signal(2, 20000 * 13, 34).reshape(2, 20000, 13), True)
from speechsauce_b200 import hostmem  # noqa: E402
xp = hostmem.pinned_empty(x.shape)
xp[...] = x
sb.feature.features(xp, cfg, lmfe=True, mfcc=True, cmvn_mode=2, empty=hostmem.pinned_empty)
tiled = sb.SpeechConfig(**C1, devices=[0]).set_tuning(tile_frames=150, chunk_bytes=1 << 20)
long_x = o.synth_signal(2, 16000 * 6 + 7, 35)
for mode in (0, 2):
    sb.feature.features(long_x, tiled, mfe=True, energy=True, lmfe=True, ssc=True, mfcc=True, cmvn_mode=mode)
sb.device.features(dx[:40], cfg, lmfe=True, mfcc=True, cmvn_mode=1)
sb.device.mel_spectrogram2(torch.from_numpy(o.synth_signal(2, 4001, 36)).cuda(), k4, log=True)
# fft_points = 512 on the STFT / librosa paths (k_stft512 + k_mel_from_power): vector and per-sample frames, odd frame counts
for H in (256, 160, 101):
    k5 = sb.SpeechConfig(16000, 512, (H + 0.5) / 16000, 0.01, 13, 40, 0.0, 8000.0, devices=[0])
    y5 = o.synth_signal(3, 4001, 37)
    sb.functions.stft2(y5, k5)
    sb.feature.mel_spectrogram2(y5, k5, log=True)
sb.feature.melspectrogram_librosa(o.synth_signal(2, 5000, 38), sb.feature.librosa_config(16000, 512, 160, 80, devices=[0]), db=True)
# librosa mode at n_fft = 2048 (k_stft2048<2> + k_mel_from_power<16>), the pipelined k_stft2048m with several rounds per
# group (the filterbank product of round r inside round r + 1), the register-resident and the cooperative cmvn
sb.feature.melspectrogram_librosa(o.synth_signal(2, 9000, 39), sb.feature.librosa_config(22050, 2048, 512, 128, devices=[0]), db=True)
sb.feature.melspectrogram_librosa(o.synth_signal(1, 2048, 40), sb.feature.librosa_config(16000, 2048, 250, 40, devices=[0]))
sb.feature.mel_spectrogram2(o.synth_signal(150, 22050, 41), sb.SpeechConfig(**C3, devices=[0]), log=True)
sb.processing.cmvn(o.synth_signal(40, 60 * 12, 42).reshape(40, 60, 12), True)
sb.processing.cmvn(o.synth_signal(9, 2 * 20, 43).reshape(9, 2, 20), False)
sb.processing.cmvn(o.synth_signal(5, 70000 * 13, 44).reshape(5, 70000, 13), True)
sb.feature.melspectrogram_librosa(o.synth_signal(3, 3000, 45), sb.feature.librosa_config(8000, 256, 80, 40, devices=[0]), db=True)  # k_stft256
sb.functions.stft2(o.synth_signal(2, 2001, 46), sb.SpeechConfig(8000, 256, 77.5 / 8000, 0.01, 13, 20, 0.0, 4000.0, devices=[0]))
torch.cuda.synchronize()
print("sanitize_small: all kernels ran")
