#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of the built libpxr.so (the .so itself is git-ignored): the evidence that the hot
kernels are tcgen05 / TMEM / TMA-engine code and which ones use the legacy warp-level MMA.

  python scripts/sass_opcodes.py [libpxr.so] > profiles/sass_opcodes.json

Mnemonics (B200_PROFILING.md): UTCHMMA / UTCQMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit,
UBLKCP = cp.async.bulk (TMA engine, 1-D), UTMALDG = TMA tensor load, HMMA = mma.sync, SYNCS = mbarrier, USETMAXREG = setmaxnreg.
"""
import json
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
WATCH = ("UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "HMMA", "SYNCS", "USETMAXREG",
         "LDL", "STL", "LDG", "STG", "LDS", "STS", "SHFL", "FFMA", "MUFU", "ATOMS", "BAR")


def main():
    lib = Path(sys.argv[1]) if len(sys.argv) > 1 else REPO / "pixelrec_multimodal_b200" / "libpxr.so"
    out = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
    kernels, cur = {}, None
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            kernels[cur] = Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)(\.[A-Z0-9_.]+)?", ln)
        if m and cur:
            op, mods = m.group(1), m.group(2) or ""
            kernels[cur]["_total"] += 1
            if op in WATCH:
                kernels[cur][op] += 1
                if op in ("UTCHMMA", "UTCBAR", "HMMA", "UBLKCP") and mods:
                    kernels[cur][op + mods] += 1
    demangle = subprocess.run(["cu++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines() \
        if kernels else []
    rows = []
    for (name, c), pretty in zip(kernels.items(), demangle + [""] * len(kernels)):
        rows.append({"kernel": (pretty or name)[:160], "instructions": c.pop("_total", 0), "opcodes": dict(sorted(c.items()))})
    rows.sort(key=lambda r: -r["instructions"])
    total = Counter()
    for r in rows:
        for k, v in r["opcodes"].items():
            if "." not in k:
                total[k] += v
    json.dump({"library": str(lib.relative_to(REPO)) if lib.is_relative_to(REPO) else str(lib), "arch": "sm_100a",
               "n_kernels": len(rows), "library_totals": dict(sorted(total.items())), "kernels": rows}, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
