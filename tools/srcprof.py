#!/usr/bin/env python
"""srcprof.py REPORT.ncu-rep [MIN_PERCENT] - warp-stall samples and executed instructions per CUDA source line of the first kernel in an
`ncu --set full --import-source on` capture (needs -lineinfo), plus the dominant stall reason of each line."""
import csv, io, subprocess, sys

rep = sys.argv[1]
minp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur, hdr, per = None, None, {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Line No":
        hdr = r; iS = hdr.index("# Samples"); iE = hdr.index("Instructions Executed")
        st = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) < len(hdr): continue
    try: ln = int(r[0]); s = int(r[iS]); e = int(r[iE])
    except ValueError: continue
    reasons = sorted(((int(r[i] or 0), h) for i, h in st), reverse=True)[:2]
    per[(cur, ln)] = (s, e, r[1], reasons)
tot = sum(v[0] for v in per.values()) or 1
tote = sum(v[1] for v in per.values())
print(f"samples {tot}, instructions {tote}")
for (f, ln), (s, e, src, reasons) in per.items():
    if 100.0 * s / tot >= minp:
        rs = " ".join(f"{h[6:]}:{100 * n // max(s, 1)}%" for n, h in reasons if n)
        print(f"{f}:{ln:5d} {100 * s / tot:5.2f}% {100 * e / tote:5.2f}%i  [{rs}]  {src.strip()[:100]}")
