#!/usr/bin/env python3
"""Static SASS size of k_solve by source region (no GPU needed): nvdisasm line info of the in-tree library.
The solver kernel is instruction-fetch sensitive, so code size is tracked like a performance number."""
import collections, os, re, subprocess, sys, tempfile
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hkd-mpc_b200", "libhsddp_b200.so")
kern = sys.argv[2] if len(sys.argv) > 2 else "k_solve"
detail = sys.argv[3] if len(sys.argv) > 3 else None  # file name: print per-line counts for it
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(tmp) if f.startswith("hsddp_kernels.")][0]
txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
cnt = collections.Counter(); inl = collections.Counter(); cur = None; on = False
for ln in txt:
    if ln.startswith(".text."):
        on = kern in ln
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        cnt[cur] += 1
byfile = collections.Counter()
for (f, l), n in cnt.items():
    byfile[f] += n
tot = sum(cnt.values())
print(f"{kern}: {tot} instructions, {tot * 16 / 1024:.1f} KB")
for f, n in byfile.most_common():
    print(f"  {f:32s} {n:6d}  {n * 16 / 1024:6.1f} KB")
if detail:
    rows = sorted(((l, n) for (f, l), n in cnt.items() if f == detail))
    # bucket by 10 lines
    b = collections.Counter()
    for l, n in rows:
        b[l // 10 * 10] += n
    for l in sorted(b):
        print(f"    {detail}:{l:4d}-{l + 9:4d} {b[l]:6d}")
