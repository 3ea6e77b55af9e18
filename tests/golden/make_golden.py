"""Generates the golden fixtures from the oracle (the reference is a Rust crate that cannot be
built in this image and ships no golden vectors; see oracle/speechsauce_oracle.py header).
Run from the repo root:  python tests/golden/make_golden.py"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "oracle"))
import speechsauce_oracle as o  # noqa: E402

OUT = Path(__file__).parent
CASES = {
    "c1_mfcc_1s_16k": dict(kind="speechpy", n_clips=1, clip_len=16000, seed=0x5EED0001,
                           config=dict(sample_rate=16000, fft_points=512, frame_length=0.025, frame_stride=0.01,
                                       num_cepstral=13, num_filters=40, low_frequency=0.0, high_frequency=8000.0)),
    "c2_kws_slice": dict(kind="speechpy", n_clips=4, clip_len=16000, seed=0x5EED0002,
                         config=dict(sample_rate=16000, fft_points=512, frame_length=0.025, frame_stride=0.01,
                                     num_cepstral=13, num_filters=40, low_frequency=0.0, high_frequency=8000.0)),
    "c3_mel_22k": dict(kind="stft", n_clips=2, clip_len=22050, seed=0x5EED0003,
                       config=dict(sample_rate=22050, fft_points=2048, frame_length=512 / 22050, frame_stride=0.01,
                                   num_cepstral=13, num_filters=128, low_frequency=0.0, high_frequency=11025.0)),
    "c4_whisper": dict(kind="stft", n_clips=2, clip_len=16000, seed=0x5EED0004,
                       config=dict(sample_rate=16000, fft_points=400, frame_length=0.01, frame_stride=0.01,
                                   num_cepstral=13, num_filters=80, low_frequency=0.0, high_frequency=8000.0)),
}

if __name__ == "__main__":
    for name, m in CASES.items():
        cfg = o.SpeechConfig(**m["config"])
        x = o.synth_signal(m["n_clips"], m["clip_len"], m["seed"])
        if m["kind"] == "speechpy":
            np.savez_compressed(OUT / f"{name}.npz", mfcc=np.stack([o.mfcc(c, cfg) for c in x]),
                                mfe=np.stack([o.mfe(c, cfg)[0] for c in x]),
                                energy=np.stack([o.mfe(c, cfg)[1] for c in x]))
        else:
            np.savez_compressed(OUT / f"{name}.npz", mel=o.mel_spectrogram2(x, cfg))
    (OUT / "golden.json").write_text(json.dumps(CASES, indent=1))
    print("wrote", list(CASES))
