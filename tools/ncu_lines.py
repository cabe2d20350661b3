#!/usr/bin/env python3
"""Aggregate an ncu report's warp-stall samples by CUDA source line (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur = None; hdr = None; agg = {}; byfunc = collections.Counter()
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and r and r[0].isdigit():
        d = dict(zip(hdr, r))
        try: smp = int(d["# Samples"]); inst = int(d["Instructions Executed"])
        except Exception: continue
        k = (cur, int(r[0]))
        a = agg.setdefault(k, [0, 0, r[1], 0])
        a[0] += smp; a[1] += inst
        try: a[3] += int(d.get("L1 Wavefronts Shared Excessive", "0") or 0)
        except Exception: pass
tot = sum(v[0] for v in agg.values()) or 1
byfile = collections.Counter()
for (f, l), v in agg.items(): byfile[f] += v[0]
print("total samples", tot, {k: round(100 * v / tot, 1) for k, v in byfile.items()})
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*v[0]/tot:5.1f}% {v[1]:>11} inst excess_wf {v[3]:>9}  {f}:{l}  {v[2].strip()[:105]}")
