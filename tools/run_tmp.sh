#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2j_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -15 gpurun_out/r2j_gputests.log
python - <<'PY'
import sys, ctypes as C, time
sys.path[:0]=['mfcc-rust_b200','oracle']
import torch, numpy as np
import speechsauce_b200 as sb
from speechsauce_b200 import _lib
for name, kw, n, ttd in (("stft W512 H256", dict(sample_rate=16000, fft_points=512, frame_length=256.5/16000, frame_stride=0.01, num_cepstral=13, num_filters=64, low_frequency=0.0, high_frequency=8000.0), 8192, 160000),
                         ("librosa 512/160", None, 8192, 160000)):
    x = torch.empty((n, ttd), device='cuda').normal_(0, 0.1)
    for flags, tag in ((0, 'fast'), (_lib.SS_FLAG_FORCE_GENERIC, 'generic')):
        if kw:
            cfg = sb.SpeechConfig(**kw, devices=[0], flags=flags)
            f = lambda: sb.device.mel_spectrogram2(x, cfg)
        else:
            cfg = sb.feature.librosa_config(16000, 512, 160, 80, devices=[0], flags=flags)
            T = 1 + ttd // 160
            out = torch.empty((n, 80, T), device='cuda')
            f = lambda: _lib.check(sb.lib.ss_librosa_melspectrogram_dev(cfg.plan, 0, None, C.c_void_p(x.data_ptr()), n, ttd, ttd, C.c_void_p(out.data_ptr()), 0))
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(name, tag, round(ms, 2), 'ms', '%.3e audio-s/s' % (n * ttd / 16000 / ms * 1e3))
PY
