#!/usr/bin/env python
"""Join an ncu `--page source --csv` (SASS rows with stall samples) with nvdisasm -gi line info and print the
source lines (of the outermost, non-inlined file position) that collect the most stall samples.

  ncu_lines.py <src.csv> <nvdisasm.sass> <mangled-kernel-substring> [top_n]
"""
import csv, re, sys
from collections import defaultdict

src_csv, sass, sub = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# ---- nvdisasm: offset -> (file:line of the outermost position)
off2line = {}
infn = False
cur = None
for ln in open(sass):
    if ln.startswith(".text."):
        infn = sub in ln
        continue
    if not infn:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))      # last annotation before an instruction = outermost
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        off2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
base = None
agg = defaultdict(lambda: defaultdict(float))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    a = int(r[ci["Address"]], 16)
    base = a if base is None else base
    off = a - base
    line = off2line.get(off, (None, ""))[0]
    n = float(r[ci["# Samples"]] or 0)
    agg[line]["samples"] += n
    agg[line]["inst"] += float(r[ci["Instructions Executed"]] or 0)
    tot += n
    for h in stall_cols:
        agg[line][h] += float(r[ci[h]] or 0)
print(f"total samples {tot:.0f}")
for line, d in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(((h[6:], d[h]) for h in stall_cols if d[h] > 0), key=lambda x: -x[1])[:4]
    print(f"{d['samples'] / tot * 100:6.2f}%  inst {d['inst']:12.0f}  {line}  " + " ".join(f"{k}={v / max(d['samples'], 1) * 100:.0f}%" for k, v in st))
