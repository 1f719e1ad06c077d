#!/usr/bin/env python
"""Per-kernel census of the Blackwell-specific SASS opcodes in libmzb200.so (cuobjdump -sass): which kernels issue tcgen05 MMAs
(UTCHMMA), TMA tensor loads / stores (UTMALDG / UTMASTG), tensor-memory loads (LDTM), legacy warp MMAs (HMMA) ...

    python profiles/sass_census.py [path/to/libmzb200.so] > profiles/r2_sass_census.txt
"""
from __future__ import annotations

import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTMACMDFLUSH", "SYNCS", "HMMA", "LDSM",
       "LDGSTS", "UCGABAR", "ELECT", "REDG", "ATOMG", "MEMBAR", "FENCE", "CCTL", "FFMA", "LDG", "STG", "LDS", "STS")


def census(so: str):
    txt = subprocess.run(["cuobjdump", "-sass", so], check=True, capture_output=True, text=True).stdout
    per, cur = collections.OrderedDict(), None
    for ln in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", ln)
        if m and cur is not None:
            cur[m.group(1)] += 1
            cur["__total__"] += 1
            full = m.group(1) + m.group(2)
            if m.group(1) in ("UTCHMMA", "UTMALDG", "UTMASTG", "UTCBAR", "HMMA", "LDTM"):
                cur["full:" + full] += 1
    return per


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.splitlines()
        return dict(zip(names, out))
    except Exception:                            # noqa: BLE001
        return {n: n for n in names}


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "muzero-breakout_b200", "libmzb200.so")
    per = census(so)
    names = demangle(list(per))
    print(f"# SASS opcode census of {os.path.relpath(so, ROOT)} (cuobjdump -sass, sm_100a); counts are static instructions per kernel")
    print("# columns: " + " ".join(OPS))
    for fn, c in per.items():
        short = re.sub(r"\(.*", "", names[fn].replace("(anonymous namespace)::", "")).replace("void ", "")
        cols = " ".join(f"{op}={c[op]}" for op in OPS if c[op])
        print(f"{short}  [{c['__total__']} instr]  {cols}")
        variants = sorted(k[5:] for k in c if k.startswith("full:"))
        if variants:
            print("    variants: " + ", ".join(f"{v} x{c['full:' + v]}" for v in variants))
    tot = collections.Counter()
    for c in per.values():
        tot.update({k: v for k, v in c.items() if not k.startswith("full:")})
    print("# whole library: " + " ".join(f"{op}={tot[op]}" for op in OPS if tot[op]))


if __name__ == "__main__":
    main()
