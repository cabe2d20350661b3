#!/usr/bin/env python3
"""Aggregate an ncu SASS source page by CUDA source line.

  ncu -i rep.ncu-rep --page source --csv --print-source sass --launch-skip K --launch-count 1 > sass.csv
  cuobjdump -xelf all libhsddp_b200.so ; nvdisasm -g hsddp_kernels.sm_100a.cubin > dis.txt
  python tools/ncu_by_line.py sass.csv dis.txt '<mangled kernel name>' [metric ...]

Instruction i of the kernel in the ncu page is matched with instruction i of the nvdisasm listing (same order), whose
`//## File ..., line N` markers give the innermost source line.  Prints the top lines for samples, shared-memory
wavefronts and any extra metric columns named on the command line.
"""
import csv, re, sys, collections

sass_csv, dis_txt, kernel = sys.argv[1:4]
extra = sys.argv[4:]
rows = list(csv.reader(open(sass_csv)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
inst = []
for r in rows[hdr_i + 1:]:  # (ncu prints the page once per view; keep the first copy)
    if r and r[0] == "Kernel Name":
        break
    if r and r[0].startswith("0x"):
        inst.append(r)
lines = []
cur = None
on = False
for l in open(dis_txt):
    if l.startswith(".text."):
        on = l.strip() == f".text.{kernel}:"
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append(cur)
assert len(lines) == len(inst), (len(lines), len(inst))
cols = {"samples": hdr.index("# Samples"), "inst": hdr.index("Instructions Executed"), "smem_wf": hdr.index("L1 Wavefronts Shared"),
        "smem_ideal": hdr.index("L1 Wavefronts Shared Ideal"), "l2_sectors": hdr.index("L2 Theoretical Sectors Global")}
for e in extra:
    cols[e] = hdr.index(e)
agg = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
for ln, r in zip(lines, inst):
    for k, c in cols.items():
        try:
            v = float(r[c])
        except ValueError:
            v = 0.0
        agg[ln][k] += v
        tot[k] += v
srcs = {}
def src(ln):
    f, n = ln
    if f not in srcs:
        import glob
        p = glob.glob(f"hkd-mpc_b200/csrc/{f}") or glob.glob(f"hkd-mpc_b200/**/{f}", recursive=True)
        srcs[f] = open(p[0]).read().split("\n") if p else []
    return srcs[f][n - 1].strip()[:110] if 0 < n <= len(srcs[f]) else ""
for key in ["samples", "smem_wf"] + extra:
    print(f"== top lines by {key} (total {tot[key]:.0f})")
    for ln, c in sorted(agg.items(), key=lambda kv: -kv[1][key])[:28]:
        print(f"  {100 * c[key] / max(tot[key], 1):5.1f}%  {ln[0]}:{ln[1]:<5d} inst {c['inst']:>10.0f} wf {c['smem_wf']:>10.0f}/{c['smem_ideal']:<10.0f} | {src(ln)}")
