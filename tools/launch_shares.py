#!/usr/bin/env python3
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv)
into per-kernel shares of ONE solve of the phased driver: the launches between the first two k_iota launches.

  python tools/launch_shares.py profiles/r01j_launches.csv > profiles/r01j_phased_solve_launches.json
"""
import csv, json, sys, collections

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
iK, iM, iV, iU, iID = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
launches = collections.OrderedDict()
for r in rows[1:]:
    d = launches.setdefault(int(r[iID]), {"kernel": r[iK]})
    v = float(r[iV].replace(",", ""))
    if r[iM] == "gpu__time_duration.sum":
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iU], 1e-6)  # -> ms
    elif r[iU] in ("Kbyte", "Mbyte", "Gbyte"):
        v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[r[iU]]
    d[r[iM]] = v
ids = sorted(launches)
iota = [i for i in ids if "k_iota" in launches[i]["kernel"]]
assert len(iota) >= 2, "need two phased solves in the capture"
names = {"k_phase<0>": "begin", "k_phase<1>": "prep (cost + LQ)", "k_phase<2>": "backward sweep", "k_sweep_w1": "backward sweep", "k_sweep_w2": "backward sweep",
         "k_phase<3>": "forward (linear rollout + line search)"}
agg = collections.OrderedDict()
for i in ids:
    if not (iota[0] <= i < iota[1]):
        continue
    k = launches[i]["kernel"]
    name = next((v for s, v in names.items() if s in k), k.split("(")[0])
    a = agg.setdefault(name, {"launches": 0, "ms_serialised": 0.0, "dram_bytes": 0.0})
    a["launches"] += 1
    a["ms_serialised"] += launches[i].get("gpu__time_duration.sum", 0.0)
    a["dram_bytes"] += launches[i].get("dram__bytes_read.sum", 0.0) + launches[i].get("dram__bytes_write.sum", 0.0)
tot = sum(a["ms_serialised"] for a in agg.values())
for a in agg.values():
    a["share"] = a["ms_serialised"] / tot
out = {"what": "one cold solve with the phased driver (launches between the first two k_iota launches), ncu --metrics "
               "gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none: serialised, "
               "cold-cache per-launch times (compare SHARES)",
       "kernels": dict(sorted(agg.items(), key=lambda kv: -kv[1]["ms_serialised"])),
       "total_ms_serialised": tot, "total_dram_bytes": sum(a["dram_bytes"] for a in agg.values())}
json.dump(out, sys.stdout, indent=1)
print()
