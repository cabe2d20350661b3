"""The pair-row shared-memory layout of the sweep tiles (ro / zo in csrc/hsddp_device.cuh) must be free of bank
conflicts for both fragment access patterns of the FP64 m8n8k4 MMA.  Pure arithmetic on the formulas in the header
(parsed, not restated), with the shared-memory model of the profiling guide: 32 banks of 4 bytes; an 8-byte access is
served per half-warp, a 16-byte access per quarter-warp."""
import os
import re

HDR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hkd-mpc_b200", "csrc", "hsddp_device.cuh")


def _fn(name):
    src = open(HDR).read()
    m = re.search(r"constexpr int %s\(int r\) \{ return ([^;]+); \}" % name, src)
    assert m, f"{name}() not found in hsddp_device.cuh"
    expr = m.group(1)
    return lambda r: eval(expr, {"r": r})


def _banks(addr_doubles, nbytes):
    first = (addr_doubles * 2) % 32
    return {(first + i) % 32 for i in range(nbytes // 4)}


def _conflict_free(lane_addrs, nbytes):
    seen = set()
    for a in lane_addrs:
        b = _banks(a, nbytes)
        if seen & b:
            return False
        seen |= b
    return True


def test_rows_do_not_overlap():
    ro, zo = _fn("ro"), _fn("zo")
    assert all(ro(r + 1) - ro(r) >= 24 for r in range(24)) and ro(0) == 0
    assert all(zo(r + 1) - zo(r) >= 16 for r in range(24)) and zo(0) == 0
    # the additivity the kernels rely on: rows t, t+4, t+8 .. are a constant apart, even rows add up
    assert all(ro(r + 4) - ro(r) == ro(4) for r in range(20)) and all(zo(r + 4) - zo(r) == zo(4) for r in range(20))
    assert all(ro(2 * a + 2 * b) == ro(2 * a) + ro(2 * b) for a in range(6) for b in range(6))
    assert all(ro(2 * a + 1) == ro(2 * a) + 24 for a in range(12))


def test_operand_loads_are_conflict_free():
    # element (t + k, c0 + g) per lane 4 g + t, 8 bytes: one half-warp = g in 0..3 (lanes 0..15) or 4..7
    for fn, width in ((_fn("ro"), 24), (_fn("zo"), 16)):
        for k in range(0, 24, 4):
            for c0 in range(0, width, 8):
                for half in (0, 1):
                    lanes = [fn(t + k) + c0 + g for g in range(4 * half, 4 * half + 4) for t in range(4)]
                    assert _conflict_free(lanes, 8), (width, k, c0, half)


def test_accumulator_tiles_are_conflict_free():
    # element (8 I + g, c0 + 2 t .. + 1) per lane 4 g + t, 16 bytes: one quarter-warp = rows g = 2 q, 2 q + 1
    for fn, width in ((_fn("ro"), 24), (_fn("zo"), 16)):
        for I in range(3):
            for c0 in range(0, width, 8):
                for q in range(4):
                    lanes = [fn(8 * I + g) + c0 + 2 * t for g in (2 * q, 2 * q + 1) for t in range(4)]
                    assert _conflict_free(lanes, 16), (width, I, c0, q)


def test_no_uniform_stride_serves_both_patterns():
    # the reason for the pair layout: with a uniform row stride S one of the two patterns conflicts
    for S in range(24, 65):
        ops = all(_conflict_free([S * t + g for g in range(4) for t in range(4)], 8) for _ in (0,))
        acc = _conflict_free([S * g + 2 * t for g in (0, 1) for t in range(4)], 16)
        assert not (ops and acc), S
