#!/usr/bin/env python
"""Attribute an ncu capture to CUDA source lines.

usage: python profiles/ncu_lines.py <report.ncu-rep> <kernel-substring> [top_n]

ncu's CSV source page only lists SASS; this joins it (by instruction order) with `nvdisasm -g`
line info of the library that was profiled (rebuild the same commit first), then prints executed
warp instructions and stall samples per source line, per opcode, and the stall-reason mix."""
import collections
import csv
import re
import subprocess
import sys
import tempfile
from pathlib import Path

rep, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = Path(__file__).resolve().parent.parent
lib = root / "mfcc-rust_b200" / "lib" / "libspeechsauce_cuda.so"
tmp = Path(tempfile.mkdtemp())
subprocess.run(["cuobjdump", "-xelf", "all", str(lib)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
seq, srcfile = [], None
for cubin in tmp.glob("*.cubin"):
    txt = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True).stdout
    infn, cur = False, None
    for l in txt.split("\n"):
        if ".section" in l and ".text." in l:
            infn = kname in l
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            seq.append((cur, m.group(2)))
    if seq:
        break
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.split("\n")))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) >= len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
assert len(seq) == len(data), f"SASS length mismatch ({len(seq)} vs {len(data)}): rebuild the profiled commit"
byline, sline, byop, stall = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
for (loc, _), r in zip(seq, data):
    n = int(r[ix["Instructions Executed"]] or 0)
    s = int(r[ix["# Samples"]] or 0)
    byline[loc] += n
    sline[loc] += s
    toks = r[ix["Source"]].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    byop[op.split(".")[0]] += n
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h and r[ix[h]]:
            stall[h] += int(r[ix[h]])
tot, st = sum(byline.values()), max(1, sum(sline.values()))
print(f"kernel {kname}: {len(seq)} SASS instructions, {tot} warp instructions executed, {st} stall samples")
print("--- opcodes")
for op, n in byop.most_common(16):
    print(f"{op:10s} {100 * n / tot:5.1f}%")
print("--- stall reasons")
for k, n in stall.most_common(10):
    print(f"{k:26s} {100 * n / st:5.1f}%")
print("--- source lines (inst %, stall %)")
cache = {}
for loc, n in sorted(byline.items(), key=lambda x: -x[1])[:top]:
    text = ""
    if loc:
        f = cache.setdefault(loc[0], Path(loc[0]).read_text().split("\n") if Path(loc[0]).exists() else [])
        text = f[loc[1] - 1].strip()[:88] if len(f) >= loc[1] else ""
    print(f"{(Path(loc[0]).name + ':' + str(loc[1])) if loc else '?':28s} {100 * n / tot:5.1f}% {100 * sline[loc] / st:5.1f}% | {text}")
