#!/usr/bin/env python
"""Copies the round-2 bench lines from gpurun_out/ into profiles/ and writes profiles/bench_r2_scaling.md."""
import json
import shutil
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
G, P = ROOT / "gpurun_out", ROOT / "profiles"


def last_json(path):
    try:
        return json.loads(path.read_text().strip().splitlines()[-1])
    except Exception:
        return None


rows = []
for f in sorted(G.glob("r2s_c*_n*.json")):
    d = last_json(f)
    if not d:
        continue
    name = f.stem[4:]  # c2_weak_n8
    t = last_json(G / f"r2t_{name}.json") if (G / f"r2t_{name}.json").exists() else None
    if t:  # device-resident fields re-measured with the final kernels (tools/run_scaling_c34.sh)
        for k in ("ms_per_step", "value", "roofline", "clocks", "steps", "warmup", "gpu_launches", "frames_per_s"):
            if k in t:
                d[k] = t[k]
        d["remeasured"] = ("device-resident fields (ms_per_step, value, roofline, clocks) from tools/run_scaling_c34.sh on the final kernels; "
                           "the e2e fields, where present, from tools/run_scaling.sh earlier in the round (PCIe-bound)")
        (P / f"bench_r2_{name}.json").write_text(json.dumps(d) + "\n")
    else:
        shutil.copy(f, P / f"bench_r2_{name}.json")
    cfg, sc, n = name.split("_")
    rows.append((cfg, sc, int(n[1:]), d))
for f in sorted(G.glob("r2s_pcie_n*.json")):
    shutil.copy(f, P / f"pcie_r2_{f.stem.split('_')[-1]}.json")

one = {}
for c in ("c1", "c2", "c3", "c4", "c5"):
    for cand in (G / f"r2z_bench_{c}.json", G / f"r2g_bench_{c}.json", G / f"r2h_bench_{c}.json", G / "r2a_bench_c2.json"):
        d = last_json(cand) if cand.exists() else None
        if d and d.get("config", {}).get("workload", "").startswith(c):
            one[c] = d
            shutil.copy(cand, P / f"bench_r2_{c}.json")
            break

v = lambda d, k: (d.get(k) or {}).get("value")
fmt = lambda x: "—" if x is None else f"{x:.3e}"
out = ["# Round-2 bench lines: every named shape at 1 / 2 / 4 / 8 B200s", "",
       "`bench.py` under `torchrun` (one process per GPU, no data-path collective); device-resident `value` timed with CUDA events,",
       "max over ranks; `e2e` = blocking host-buffer call on pinned memory; `ceiling` = bare concurrent H2D + D2H copies of the same",
       "bytes on every rank (no kernels); `inproc` = ONE plan over all N devices inside rank 0.  audio-s/s throughout.", "",
       "| config | scaling | N | clips / GPU | ms / step | audio-s/s | vs N = 1 | e2e pinned | e2e / ceiling | e2e pageable | e2e in-process | PCIe ceiling |",
       "|---|---|---|---|---|---|---|---|---|---|---|---|"]
base = {c: d["value"] for c, d in one.items()}
for c in ("c2", "c3", "c4", "c5"):
    if c in one:
        d = one[c]
        e = d.get("e2e") or {}
        out.append(f"| {c} | — | 1 | {d['config']['clips_per_gpu']} | {d['ms_per_step']:.3f} | {d['value']:.3e} | 1.00 | {fmt(v(d, 'e2e'))} | "
                   f"{e.get('frac_of_pcie_ceiling', float('nan')):.2f} | {fmt(v(d, 'e2e_pageable'))} | — | {fmt(v(d, 'pcie_ceiling'))} |")
    for sc in ("weak", "strong"):
        for cfg, s, n, d in sorted(rows, key=lambda r: r[2]):
            if cfg != c or s != sc:
                continue
            e = d.get("e2e") or {}
            fr = e.get("frac_of_pcie_ceiling")
            out.append(f"| {c} | {sc} | {n} | {d['config']['clips_per_gpu']} | {d['ms_per_step']:.3f} | {d['value']:.3e} | "
                       f"{d['value'] / base[c]:.2f} | {fmt(v(d, 'e2e'))} | {'—' if fr is None else f'{fr:.2f}'} | {fmt(v(d, 'e2e_pageable'))} | "
                       f"{fmt(v(d, 'e2e_inproc'))} | {fmt(v(d, 'pcie_ceiling'))} |")
out += ["", "PCIe ceiling of the box on its own (`tools/pcie_ceiling.py`, the c2 byte counts: 4.19 GB in, 1.35 GB out per rank):", "",
        "| N | both directions, GB/s per GPU | total GB/s | H2D alone per GPU | D2H alone per GPU |", "|---|---|---|---|---|"]
for n in (1, 2, 4, 8):
    for cand in (G / f"r2s_pcie_n{n}.json", G / "r2g_pcie_1gpu.json"):
        d = last_json(cand) if cand.exists() else None
        if d and d.get("n_gpus") == n:
            out.append(f"| {n} | {d['both_gb_s_per_gpu']:.1f} | {d['both_gb_s_total']:.1f} | {d['h2d_gb_s_per_gpu']:.1f} | {d['d2h_gb_s_per_gpu']:.1f} |")
            if n == 1 and d.get("host_register"):
                r = d["host_register"]
                reg = (f"cudaHostRegister {r['register_gb_s']:.1f} GB/s, cudaHostUnregister {r['unregister_gb_s']:.1f} GB/s, H2D from the registered range "
                       f"{r['h2d_registered_gb_s']:.1f} GB/s, one numpy copy thread {r['numpy_copy_1_thread_gb_s']:.1f} GB/s")
            break
out += ["", f"Host-side rates on the same box (1 GiB pageable buffer): {reg}."]
(P / "bench_r2_scaling.md").write_text("\n".join(out) + "\n")
print("\n".join(out))
