#!/usr/bin/env python3
"""profiles/k_solve_traffic.json (what bench.py reports as roofline.traffic) from an ncu launch list of bench.py:

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file launches.csv python bench.py --steps 2 --warmup 1 --no-latency --no-cpu-baseline
  python tools/make_traffic_json.py launches.csv profiles/<name>_launches.csv > profiles/k_solve_traffic.json

One STEP of the bench = the launches from one cold-start reset (k_step) up to the next: the LAST complete step of the
capture is aggregated (DRAM bytes read / written, serialised kernel time, per-kernel shares)."""
import csv, json, sys, collections

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
iK, iM, iV, iU, iID = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
launches = collections.OrderedDict()
for r in rows[1:]:
    d = launches.setdefault(int(r[iID]), {"kernel": r[iK]})
    v = float(r[iV].replace(",", ""))
    if r[iM] == "gpu__time_duration.sum":
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iU], 1e-6)
    elif r[iU] in ("Kbyte", "Mbyte", "Gbyte"):
        v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[r[iU]]
    d[r[iM]] = v
ids = sorted(launches)
resets = [i for i in ids if "k_step" in launches[i]["kernel"]]
# the bench's steps are reset + solve; the last two resets bracket the last complete step that is followed by another reset
steps = [(a, b) for a, b in zip(resets, resets[1:]) if sum("k_iota" in launches[i]["kernel"] or "k_solve" in launches[i]["kernel"] for i in ids if a <= i < b) >= 1]
a, b = steps[-1]
names = {"k_phase<0>": "begin", "k_phase<1>": "prep (cost + LQ)", "k_phase<2>": "backward sweep (four warps per problem)", "k_sweep_w1": "backward sweep (one warp per problem)",
         "k_phase<3>": "forward (line search; the linear rollout too when k_lr_w1 is off)", "k_lr_w1": "linear rollout (one warp per problem)", "k_step": "cold-start reset"}
agg = collections.OrderedDict()
rd = wr = ms = 0.0
for i in ids:
    if not (a <= i < b):
        continue
    k = launches[i]["kernel"]
    name = next((v for s, v in names.items() if s in k), k.split("(")[0])
    q = agg.setdefault(name, {"launches": 0, "ms_serialised": 0.0, "dram_bytes": 0.0})
    q["launches"] += 1
    q["ms_serialised"] += launches[i].get("gpu__time_duration.sum", 0.0)
    q["dram_bytes"] += launches[i].get("dram__bytes_read.sum", 0.0) + launches[i].get("dram__bytes_write.sum", 0.0)
    rd += launches[i].get("dram__bytes_read.sum", 0.0); wr += launches[i].get("dram__bytes_write.sum", 0.0); ms += launches[i].get("gpu__time_duration.sum", 0.0)
for q in agg.values():
    q["share"] = q["ms_serialised"] / ms
out = {"config": "config3", "problems": 16384, "plan": 0.6, "dram_bytes_read": rd, "dram_bytes_write": wr, "kernel_ms_under_ncu": ms,
       "source": f"{sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]} (ncu per-launch dram__bytes over every launch of the last complete bench step; "
                 "per-launch times under ncu are serialised and cold-cache: compare SHARES)",
       "kernel_shares": {k: {"share": round(v["share"], 4), "launches": v["launches"], "ms_serialised": round(v["ms_serialised"], 2), "dram_gb": round(v["dram_bytes"] / 1e9, 1)}
                         for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms_serialised"])}}
json.dump(out, sys.stdout, indent=1)
print()
