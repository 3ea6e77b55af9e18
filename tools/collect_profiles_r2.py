#!/usr/bin/env python
"""Turns the raw output of tools/measure_r2.sh (gpurun_out/r2z_*, prof_*_r2.ncu-rep) into the committed evidence under
profiles/: bench lines, the condensed ncu launch list, the per-kernel ncu text summaries and ncu_traffic.json."""
import collections
import csv
import json
import re
import subprocess
import sys
from pathlib import Path

root = Path(__file__).resolve().parent.parent
g, pr = root / "gpurun_out", root / "profiles"


def last_json_line(path):
    return [l for l in path.read_text().splitlines() if l.startswith("{")][-1]


names = ["c1", "c2", "c3", "c4", "c5", "c2_reference", "c2_generic", "p_preemphasis", "p_stack", "p_cmvn", "p_cmvn_long", "p_pspec",
         "p_pspec_generic", "p_cmvn_long_3pass", "p_cmvn_smem"]
for f in names:
    src = g / f"r2z_bench_{f}.json"
    if src.exists():
        (pr / f"bench_r2_{f}.json").write_text(last_json_line(src) + "\n")

rows = [r for r in csv.reader((g / "launches_r2.csv").read_text().splitlines()) if len(r) > 10]
ix = {h: i for i, h in enumerate(rows[0])}
out = [["id", "kernel", "grid", "block", "gpu__time_duration_ns"]]
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    k = r[ix["Kernel Name"]]
    m = re.search(r"(k_[a-z0-9_]+)(<[^>]*>)?", k)
    short = (m.group(1) + (m.group(2) or "")) if m else "torch::" + ("normal_" if "normal" in k else "other")
    out.append([r[ix["ID"]], short, r[ix["Grid Size"]], r[ix["Block Size"]], r[ix["Metric Value"]]])
with open(pr / "launches_r2.csv", "w", newline="") as fh:
    csv.writer(fh).writerows(out)
tot, cnt = collections.Counter(), collections.Counter()
for r in out[1:]:
    tot[r[1]] += float(r[4])
    cnt[r[1]] += 1
for k, v in tot.most_common(8):
    print(f"launch list: {k:28s} x{cnt[k]:4d} {v / 1e6:9.3f} ms")


def run(script, *args):
    return subprocess.run([sys.executable, str(pr / script), *args], capture_output=True, text=True).stdout


def traffic(txt):
    t = 0.0
    for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        v, u = re.search(key + r"\s+([\d.]+) (G|M)byte", txt).groups()
        t += float(v) * (1e9 if u == "G" else 1e6)
    return t


tr = {}
for rep, kn, name, cfg, clips, phases in [
        ("prof_fast512c_r2", "k_fast512cILi13ELb1ELb0ELi5", "ncu_fast512c_r2", "c2", 65536, True),
        ("prof_fast512c_c5_r2", "k_fast512cILi13ELb1ELb1ELi5", "ncu_fast512c_c5_r2", "c5", 128, True),
        ("prof_stft400m_r2", "k_stft400mILb0ELb0", "ncu_stft400m_r2", None, None, False),
        ("prof_stft2048m_r2", "k_stft2048m", "ncu_stft2048m_r2", "c3", 4096, False),
        ("prof_pspec512_r2", "k_pspec512ILi13", "ncu_pspec512_r2", None, None, False),
        ("prof_cmvn_reg_r2", "k_cmvn_reg", "ncu_cmvn_reg_r2", None, None, False),
        ("prof_cmvn_coop_r2", "k_cmvn_coop", "ncu_cmvn_coop_r2", None, None, False)]:
    rp = g / f"{rep}.ncu-rep"
    if not rp.exists():
        print("missing", rp)
        continue
    txt = run("ncu_summary.py", str(rp)) + "\n" + (run("ncu_phases.py", str(rp), kn) + "\n" if phases else "") + run("ncu_lines.py", str(rp), kn, "30")
    (pr / f"{name}.txt").write_text(txt)
    if cfg:
        tr[cfg] = {"kernel": kn.split("IL")[0].split("E")[0] if kn.startswith("k_fast") else kn.split("IL")[0], "clips": clips, "dram_bytes": traffic(txt),
                   "source": f"profiles/{name}.txt (dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full)"}
(pr / "ncu_traffic.json").write_text(json.dumps(tr, indent=1) + "\n")
if (g / "r2z_time_stft.txt").exists():
    (pr / "stft_paths_r2.txt").write_text((g / "r2z_time_stft.txt").read_text())
for tool in ("memcheck", "racecheck"):
    f = g / f"r2z_sanitize_{tool}.log"
    if f.exists():
        keep = [l for l in f.read_text().splitlines() if re.search(r"SUMMARY|sanitize_small|Race reported|Invalid|hazard", l)]
        (pr / f"sanitize_r2_{tool}.log").write_text("\n".join(keep[-12:]) + "\n")
for c in ["c1", "c2", "c3", "c4", "c5"]:
    d = json.loads((pr / f"bench_r2_{c}.json").read_text())
    e = d.get("e2e") or {}
    print(c, d["kernel"], f"{d['ms_per_step']:.3f} ms", f"hbm {100 * d['roofline']['frac']:.1f} %", f"fp32 {100 * d['roofline']['fp32']['frac']:.1f} %",
          f"value {d['value']:.3e}", f"frames/s {d['frames_per_s']:.3e}", f"e2e {e.get('value')} ({e.get('ms_per_step')} ms, {e.get('frac_of_pcie_ceiling')})",
          f"pageable {(d.get('e2e_pageable') or {}).get('value')}", f"cpu {(d.get('cpu_baseline') or {}).get('value')}")
