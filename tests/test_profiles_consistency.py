"""The summaries under profiles/ that bench.py and DESIGN.md quote must follow from the committed raw ncu launch list."""
import json
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def test_launch_shares_reproduce_from_the_raw_launch_list():
    raw = os.path.join(ROOT, "profiles", "r01j_launches.csv")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_shares.py"), raw], capture_output=True, text=True, check=True).stdout
    got = json.loads(out)
    want = json.load(open(os.path.join(ROOT, "profiles", "r01j_phased_solve_launches.json")))
    assert got["kernels"].keys() == want["kernels"].keys()
    for k, v in want["kernels"].items():
        assert got["kernels"][k]["launches"] == v["launches"]
        assert abs(got["kernels"][k]["ms_serialised"] - v["ms_serialised"]) < 1e-9
        assert abs(got["kernels"][k]["dram_bytes"] - v["dram_bytes"]) < 1.0
    # one phased solve = 3 launches per group and round + one begin launch per group (+ reset and iota)
    k = got["kernels"]
    assert k["backward sweep"]["launches"] == k["prep (cost + LQ)"]["launches"] == k["forward (linear rollout + line search)"]["launches"]
    assert abs(sum(v["share"] for v in k.values()) - 1.0) < 1e-12


def test_traffic_figure_reproduces_from_the_launch_list_it_names():
    """profiles/k_solve_traffic.json (bench.py's roofline.traffic and kernel_shares) follows from the committed raw ncu
    launch list named in its `source` field (tools/make_traffic_json.py)."""
    tr = json.load(open(os.path.join(ROOT, "profiles", "k_solve_traffic.json")))
    raw = os.path.join(ROOT, tr["source"].split(" ")[0])
    assert os.path.exists(raw), raw
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_traffic_json.py"), raw, tr["source"].split(" ")[0]],
                         capture_output=True, text=True, check=True).stdout
    got = json.loads(out)
    assert abs(got["dram_bytes_read"] - tr["dram_bytes_read"]) < 1.0 and abs(got["dram_bytes_write"] - tr["dram_bytes_write"]) < 1.0
    assert got["kernel_shares"] == tr["kernel_shares"]
    assert abs(sum(v["share"] for v in got["kernel_shares"].values()) - 1.0) < 1e-3
