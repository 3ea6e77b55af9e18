#!/usr/bin/env python
"""Phase breakdown of an ncu capture of k_fast512b: share of stall samples (~time), executed warp
instructions and shared-memory wavefronts per kernel phase.
usage: python profiles/ncu_phases.py <report.ncu-rep> [kernel-substring]"""
import collections
import csv
import re
import subprocess
import sys
import tempfile
from pathlib import Path

root = Path(__file__).resolve().parent.parent
rep = sys.argv[1]
kname = sys.argv[2] if len(sys.argv) > 2 else "k_fast512bILi13ELb1ELb0"
gen = "fast512c" if "fast512c" in kname else "fast512b"
srcfile = root / "mfcc-rust_b200" / "csrc" / f"ss_kernels_{gen}.cu"
tmp = Path(tempfile.mkdtemp())
subprocess.run(["cuobjdump", "-xelf", "all", str(root / "mfcc-rust_b200" / "lib" / "libspeechsauce_cuda.so")], cwd=tmp,
               stdout=subprocess.DEVNULL, check=True)
seq = []
for cubin in tmp.glob(f"*{gen}*.cubin"):
    txt = subprocess.run(["nvdisasm", "-gi", "-c", str(cubin)], capture_output=True, text=True).stdout
    infn, cur = False, None
    for l in txt.split("\n"):
        if ".section" in l and ".text." in l:
            infn = kname in l
        if not infn:
            continue
        if "//## File" in l:
            # innermost location first, then the "inlined at" chain: keep the outermost kernel-file line
            locs = re.findall(r'"([^"]+)", line (\d+)', l)
            cur = None
            for f, ln in locs:
                if Path(f).name == srcfile.name:
                    cur = (srcfile.name, int(ln))
            if cur is None and locs:
                cur = (Path(locs[0][0]).name, int(locs[0][1]))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            seq.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) >= len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
assert len(seq) == len(data), (len(seq), len(data))
src = srcfile.read_text().split("\n")
find = lambda s: next((i + 1 for i, l in enumerate(src) if s in l), None)  # noqa: E731
marks = [("loads", find("// ---- framing: v[j]")), ("fft1", find("// ---- stage 1: 16-point")),
         ("twiddle+transpose", find("// ---- twiddle W256")), ("fft2", find("// ---- stage 2: 16-point")),
         ("split+magnitude", find("// ---- real-FFT split")), ("mel", find("// ---- banded mel reduction")),
         ("mel", find("// ---- mel filterbank on the tensor cores")), ("zh+ln+stores", find("// ---- zero_handling -> stores")),
         ("dct", find("// ---- DCT-II on the tensor cores")),
         ("stores", find("// ---- stores: [clip]")), ("dct", find("// ---- ln -> DCT-II")),
         ("cmvn", find("// ---- per-clip CMVN")), ("end", 10 ** 9)]
marks = [m for m in marks if m[1] is not None]
t, n, w = collections.Counter(), collections.Counter(), collections.Counter()
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
st = collections.defaultdict(collections.Counter)
lastk = None
for loc, r in zip(seq, data):
    if loc and loc[0] == srcfile.name:
        lastk = loc[1]  # inlined helpers (fft_ct.cuh) are attributed to the phase that calls them
    name = "prologue"
    for (nm, l0), (_, l1) in zip(marks, marks[1:]):
        if l0 <= (lastk or 0) < l1:
            name = nm
    t[name] += int(r[ix["# Samples"]] or 0)
    n[name] += int(r[ix["Instructions Executed"]] or 0)
    w[name] += int(r[ix["L1 Wavefronts Shared"]] or 0)
    for c in stall_cols:
        st[name][c[6:]] += int(r[ix[c]] or 0)
ts, tn, tw = sum(t.values()), sum(n.values()), sum(w.values())
print(f"{kname}: {tn} warp instructions, {tw} shared-memory wavefronts")
print(f"{'phase':20s} {'time':>7s} {'instr':>7s} {'smem wf':>8s}")
for nm in ["prologue"] + list(dict.fromkeys(m[0] for m in marks[:-1])):
    top = ", ".join(f"{k} {100 * v / max(1, sum(st[nm].values())):.0f}%" for k, v in st[nm].most_common(4))
    print(f"{nm:20s} {100 * t[nm] / ts:6.1f}% {100 * n[nm] / tn:6.1f}% {100 * w[nm] / tw:7.1f}%   {top}")
