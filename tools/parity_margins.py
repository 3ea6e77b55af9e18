#!/usr/bin/env python
"""Measured parity margins on a GPU: the worst error of every quantity class against the f32 oracle and the f64
oracle, next to the f32-vs-f64 distance of the oracle itself (its own rounding).  Writes one JSON object; the
table in BASELINE.md section 6 is made from it (profiles/parity_margins_r2.json)."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / "mfcc-rust_b200", ROOT / "oracle", ROOT / "tests"):
    sys.path.insert(0, str(p))
import speechsauce_b200 as sb  # noqa: E402
import speechsauce_oracle as o  # noqa: E402
from test_host_abi import CONFIGS  # noqa: E402


def rel(got, want, floor_frac):
    """max |got - want| / (|want| + floor), floor = floor_frac * max|want| (0: purely relative)"""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return float((np.abs(got - want) / (np.abs(want) + floor_frac * np.abs(want).max() + 1e-300)).max())


def absmax(got, want):
    return float(np.abs(np.asarray(got, np.float64) - np.asarray(want, np.float64)).max())


out = {}
n = 48
x = o.synth_signal(n, 16000, 4242)
x[5] *= 1e-3          # a quiet clip
x[6, 4000:] = 0.0     # silence after sound
t = np.arange(16000) / 16000.0
x[7] = (0.5 * np.sin(2 * np.pi * 440 * t) + 0.01 * x[7]).astype(np.float32)  # tonal
kw = CONFIGS["c1"]
cfg = sb.SpeechConfig(**kw, devices=[0])
oc = o.SpeechConfig(**kw)
r = sb.feature.features(x, cfg, mfe=True, energy=True, lmfe=True, ssc=True, mfcc=True)
rc = sb.feature.features(x, cfg, mfcc=True, cmvn_mode=2)
acc = {k: {"vs_f32": 0.0, "vs_f64": 0.0, "oracle_f32_vs_f64": 0.0} for k in
       ("mfe_rel", "mfe_rel_floor2e-6", "energy_rel", "ssc_rel", "lmfe_abs", "mfcc_abs", "cmvn_mfcc_abs")}


def upd(key, got, w32, w64, fn):
    a = acc[key]
    a["vs_f32"] = max(a["vs_f32"], fn(got, w32))
    a["vs_f64"] = max(a["vs_f64"], fn(got, w64))
    a["oracle_f32_vs_f64"] = max(a["oracle_f32_vs_f64"], fn(w32, w64))


for i in range(n):
    f32, e32 = o.mfe(x[i], oc)
    f64, e64 = o.mfe(x[i], oc, np.float64)
    upd("mfe_rel", r["mfe"][i], f32, f64, lambda g, w: rel(g, w, 0.0))
    upd("mfe_rel_floor2e-6", r["mfe"][i], f32, f64, lambda g, w: rel(g, w, 2e-6))
    upd("energy_rel", r["energy"][i], e32, e64, lambda g, w: rel(g, w, 0.0))
    upd("ssc_rel", r["ssc"][i], o.ssc(x[i], oc), o.ssc(x[i], oc, np.float64), lambda g, w: rel(g, w, 0.0))
    upd("lmfe_abs", r["lmfe"][i], np.log(f32), np.log(f64), absmax)
    m32, m64 = o.mfcc(x[i], oc), o.mfcc(x[i], oc, np.float64)
    upd("mfcc_abs", r["mfcc"][i], m32, m64, absmax)
    upd("cmvn_mfcc_abs", rc["mfcc"][i], o.cmvn(m32, True), o.cmvn(m64, True, np.float64), absmax)
out["speechpy_c1_48_clips"] = acc

# STFT family: c3 and c4 shapes, a few rows
for cname, ttd in (("c3", 22050 * 2), ("c4", 16000 * 3)):
    kw = CONFIGS[cname]
    cfg = sb.SpeechConfig(**kw, devices=[0])
    fresh = lambda: o.SpeechConfig(**kw)  # the oracle keeps the STFT overlap memory in its config, like the reference
    y = o.synth_signal(6, ttd, 99)
    y[2] = (0.5 * np.sin(2 * np.pi * 1000 * np.arange(ttd) / kw["sample_rate"])).astype(np.float32)
    got = sb.feature.mel_spectrogram2(y, cfg)
    w32 = o.mel_spectrogram2(y, fresh())
    w64 = o.mel_spectrogram2(y, fresh(), np.float64)
    st = sb.functions.stft2(y, sb.SpeechConfig(**kw, devices=[0]))
    s32, s64 = o.stft2(y, fresh()), o.stft2(y, fresh(), np.float64)
    d = {"mel_rel": {}, "mel_rel_floor2e-6": {}, "stft_abs_over_peak": {}}
    for key, fl in (("mel_rel", 0.0), ("mel_rel_floor2e-6", 2e-6)):
        nz = w64 > 0
        d[key] = {"vs_f32": rel(got[nz], w32[nz], fl), "vs_f64": rel(got[nz], w64[nz], fl), "oracle_f32_vs_f64": rel(w32[nz], w64[nz], fl)}
    pk = np.abs(s64).max()
    d["stft_abs_over_peak"] = {"vs_f32": float(np.abs(st - s32).max() / pk), "vs_f64": float(np.abs(st - s64).max() / pk),
                               "oracle_f32_vs_f64": float(np.abs(s32 - s64).max() / pk)}
    out[f"stft_{cname}"] = d
print(json.dumps(out, indent=1))
