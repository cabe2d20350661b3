#!/usr/bin/env python3
"""bench.py — batched HS-DDP solves/s on B200 (BASELINE.json metric).

A "step" is one cold solve of the whole batch: reset to the reference's cold-start
guess + MultiPhaseDDP::solve for every problem of the rank's shard.

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA path through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the reference's HS-DDP path on host cores

Workload: SURVEY.md §8(d) config 3 — 16,384 mixed-gait Mini Cheetah problems
(trot / bound / pronk with flight phases and reset maps), plan 0.6 s, dt 0.01, ReB+AL,
per GPU ("weak" scaling: problems are sharded by global index, no data-path collective).
Timing: CUDA events on the solver handle's stream, barrier + synchronise on both sides,
MAX over ranks.  The per-GPU workspace (~8 GB) is far larger than L2, so no explicit L2
flush is needed between timed steps.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F_STAGE = 203904.0           # algorithmic FLOP of one dense 24x24 Riccati stage
F_ITER_STAGE = 3600.0 + 5000.0   # LQ approximation + linear rollout, per stage per DDP iteration
F_TRIAL_STAGE = 2100.0       # one line-search trial, per stage
BYTES_STAGE = 6400.0         # minimum HBM bytes per stage per iteration (fused), SURVEY.md §8d


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--problems", type=int, default=16384, help="problems per GPU")
    ap.add_argument("--config", default="config3", choices=["config2", "config3", "config4"])
    ap.add_argument("--plan", type=float, default=0.6)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target wall time of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def build_workload(pkg, wl, args, rank, world):
    n = args.problems
    if args.config == "config2":
        return wl.config2(pkg, n, args.plan)
    if args.config == "config4":
        return wl.config4(pkg, n, args.plan)
    return wl.config3(pkg, n, args.plan, first=rank * n)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(smax)) if smax else None,
                "power_w_max": float(max(power)) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_run(wl_mod, pkg, w, seconds, steps=1, warmup=0):
    """Times the CPU implementation of the path (the oracle restatement running on the reference's own
    compiled CasADi model when oracle/_ref travelled, else on its port), one problem per std::thread,
    on a bounded sample of the SAME workload.  The only place bench.py executes oracle/."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as orc
    cores = orc.hardware_concurrency()
    tables = {}
    for g in wl_mod.GAITS:
        tables[g] = orc.GaitTable(wl_mod.gait_path(g))

    def run(count):
        idx = np.arange(count) % w.n
        tabs = [tables[w.keys[w.schedule_id[i]][0]] for i in idx]
        k0 = [w.keys[w.schedule_id[i]][1] for i in idx]
        wall, summ = orc.batch_solve(tabs, k0, w.x0[idx], plan=w.plan, n_threads=cores)
        return wall, summ
    pilot = max(2 * cores, 16)
    wall, _ = run(pilot)
    rate = pilot / wall
    count = int(max(pilot, min(w.n, rate * seconds)))
    count = max(cores, (count // cores) * cores)
    walls = []
    for _ in range(warmup):
        run(count)
    summ = None
    for _ in range(max(1, steps)):
        wall, summ = run(count)
        walls.append(wall)
    value = count * len(walls) / sum(walls)
    kind = "reference-model+port-solver" if orc.ref_available() else "port"
    return dict(value=value, unit="solves/s", cores=cores, kind="port",
                model="reference CasADi C compiled unmodified (oracle/_ref)" if orc.ref_available() else "oracle model port",
                sample=f"first {count} problems of the workload, one problem per std::thread on {cores} threads, "
                       f"{len(walls)} pass(es), {sum(walls):.1f} s; mean iterations {summ[:, 1].mean():.2f}",
                ms_per_step=1e3 * sum(walls) / len(walls), count=count, detail=kind)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    pkg = importlib.import_module("hkd-mpc_b200")
    wl = importlib.import_module("hkd-mpc_b200.workloads")
    sh = importlib.import_module("hkd-mpc_b200.sharding")

    if args.impl == "reference":
        # CPU arm: rank 0 alone runs; other ranks exit 0 without work
        if rank != 0:
            return 0
        w = build_workload(pkg, wl, args, 0, 1)
        r = cpu_reference_run(wl, pkg, w, args.cpu_seconds, steps=args.steps, warmup=min(args.warmup, 1))
        line = {"impl": "reference", "metric": "batched HS-DDP solves/sec", "value": r["value"], "unit": "solves/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": w.name, "plan_duration_s": args.plan, "dt": 0.01, "options": "ddp_setting.info as consumed (ReB+AL, MS)"},
                "cpu_baseline": {"value": r["value"], "unit": "solves/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"], "model": r["model"]},
                "e2e": {"value": r["value"], "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the solver has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # keep NCCL's version banner off stdout (rank 0 prints ONE JSON line); HSDDP_NCCL_DEBUG overrides
        os.environ["NCCL_DEBUG"] = os.environ.get("HSDDP_NCCL_DEBUG", "WARN")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    w = build_workload(pkg, wl, args, rank, world)
    B = pkg.MultiPhaseDDPBatch(local_rank)
    B.set_problems(w.schedules, w.schedule_id)
    opt = pkg.Options()
    # pinned host staging buffers for the end-to-end leg
    n_cmd = 8  # controls / gains shipped per MPC update (HKDMPC.cpp:245-248)
    x0_pin = torch.from_numpy(w.x0.copy()).pin_memory()
    cmd_pin = torch.zeros((w.n, pkg.CMD_DTYPE.itemsize), dtype=torch.uint8).pin_memory()
    cmd_view = cmd_pin.numpy().view(pkg.CMD_DTYPE).reshape(w.n)
    h2d = w.x0.nbytes
    d2h = cmd_pin.numel() + w.n * pkg.INFO_DTYPE.itemsize

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        B.reset()
        B.solve_async(opt)

    def step_e2e():
        B.set_initial_condition(x0_pin.numpy())
        B.reset()
        B.solve_async(opt)
        B.mpc_command(n_cmd, out=cmd_view)
        return B.info()

    # ---- kernel-resident timing ----
    B.set_initial_condition(w.x0)
    for _ in range(args.warmup):
        step_resident()
    B.sync()
    B.reset_counters()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    B.event_record(0)
    kernel_ms = 0.0
    for _ in range(args.steps):
        step_resident()
    B.event_record(1)
    B.sync()
    barrier()
    clocks = sampler.stop()
    ms_total = B.event_elapsed_ms(0, 1)
    ms_total = sh.max_over_ranks(ms_total, dev)
    info = B.info()
    cnt = B.counters()
    launches = cnt["solve_launches"] + cnt["step_launches"]
    # the dominant kernel: k_solve; its mean launch duration from the handle's own events
    B.reset(); B.event_record(2); B.solve_async(opt); B.event_record(3); B.sync()
    kernel_ms = B.last_solve_ms()
    cnt1 = B.counters()

    # ---- end-to-end timing through the public API with host buffers ----
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    B.event_record(4)
    for _ in range(args.steps):
        info_e2e = step_e2e()
    B.event_record(5)
    B.sync()
    barrier()
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = max(B.event_elapsed_ms(4, 5), e2e_wall_ms)
    e2e_ms = sh.max_over_ranks(e2e_ms, dev)

    # ---- single-solve latency (BASELINE.json: "single-solve p50 latency"): batch = 1, config 1, 101 cold solves ----
    latency = None
    if rank == 0:
        w1 = wl.config1(pkg, args.plan)
        B1 = pkg.MultiPhaseDDPBatch(local_rank)
        B1.set_problems(w1.schedules, w1.schedule_id)
        B1.set_initial_condition(w1.x0)
        dev_ms, wall_ms = [], []
        for rep in range(104):
            t1 = time.perf_counter()
            B1.reset()
            B1.solve(opt)
            i1 = B1.info()
            t2 = time.perf_counter()
            if rep >= 3:
                dev_ms.append(B1.last_solve_ms()); wall_ms.append((t2 - t1) * 1e3)
        # the MPC tick of the reference: warm re-solve from the previous solution with 2 AL x 1 DDP iterations (HKDMPC.cpp:102-103)
        tick_opt = pkg.Options(max_AL_iter=2, max_DDP_iter=1)
        tick_ms = []
        for rep in range(104):
            B1.solve(tick_opt)
            if rep >= 3:
                tick_ms.append(B1.last_solve_ms())
        latency = {"p50_ms": float(np.median(dev_ms)), "p50_wall_ms": float(np.median(wall_ms)), "p99_ms": float(np.percentile(dev_ms, 99)),
                   "mpc_tick_p50_ms": float(np.median(tick_ms)),
                   "reps": len(dev_ms), "iterations": int(i1["n_iter"][0]),
                   "what": "one cold Mini Cheetah trot solve (config 1), batch 1: p50_ms = solve kernel on the device (CUDA events), "
                           "p50_wall_ms = host wall clock of reset + solve + info read-back; mpc_tick_p50_ms = warm re-solve with 2 AL x 1 DDP iterations (the reference's MPC update, HKDMPC.cpp:102-103), device time"}
        del B1

    # ---- statistics (the only inter-GPU traffic: a few numbers per rank) ----
    per_launch_sweep_stages = (cnt1["sweep_stages"] - cnt["sweep_stages"])
    g = sh.gather_stats(sh.local_stats(info, per_launch_sweep_stages), dev)
    tot = sh.reduce_stats(g)
    n_total = int(tot["n_problems"])
    value = n_total * args.steps / (ms_total * 1e-3)
    e2e_value = n_total * args.steps / (e2e_ms * 1e-3)

    if rank == 0:
        mean_stages = np.mean([w.schedules[s].n_stages for s in w.schedule_id])
        flop_launch = (per_launch_sweep_stages * F_STAGE + float(info["n_iter"].sum()) * mean_stages * F_ITER_STAGE
                       + float(info["n_trials"].sum()) * mean_stages * F_TRIAL_STAGE)
        peak_dfma = pkg.fp64_peak_tflops(local_rank, 0)
        peak_dmma = pkg.fp64_peak_tflops(local_rank, 1)
        peak = max(peak_dfma, peak_dmma)
        achieved = flop_launch / (kernel_ms * 1e-3) / 1e12
        bytes_launch = float(info["n_iter"].sum()) * mean_stages * BYTES_STAGE
        hbm_peak = 6553.6
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
            hbm_src = "MEASURED_PEAKS.json"
        except Exception:
            hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
        hbm_ach = bytes_launch / (kernel_ms * 1e-3) / 1e9
        # DRAM bytes of one k_solve launch from the committed ncu --set full capture of this workload (profiles/)
        traffic, traffic_src = None, None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "k_solve_traffic.json")))
            if tr.get("config") == args.config and int(tr.get("problems", 0)) == w.n and abs(tr.get("plan", 0.6) - args.plan) < 1e-9:
                traffic, traffic_src = float(tr["dram_bytes_read"]) + float(tr["dram_bytes_write"]), tr.get("source")
        except Exception:
            pass
        line = {
            "metric": "batched HS-DDP solves/sec", "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w.name, "problems_per_gpu": w.n, "problems_total": n_total, "schedules_per_gpu": len(w.schedules),
                       "plan_duration_s": args.plan, "dt": 0.01, "stages_mean": float(mean_stages),
                       "options": "ddp_setting.info as consumed (alpha .1, gamma .01, 5 AL x 10 DDP, ReB+AL, MS)",
                       "cache": "inputs larger than L2: %.1f GB workspace per GPU, no flush" % (w.n * 0.5e6 / 1e9),
                       "step": "cold-start reset + solve of every problem"},
            "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms / args.steps,
                    "what": "set_initial_condition(x0 from pinned host) + reset + solve + copy-out to pinned host of the result record and the MPC command of every problem (hkd_command_lcmt payload: 8 controls, body states, 12x12 feedback blocks, foot placements; HKDMPC.cpp:207-298)"},
            "roofline": {"bound": "tensor", "pipe": "FP64 (DFMA / DMMA m8n8k4)", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes per solve (dram__bytes_read.sum + dram__bytes_write.sum over its launches)",
                         "traffic_source": traffic_src + "; captured before the trial rollouts stopped re-reading the gains, which removes about 138 KB per trial (~88 GB per solve of this workload) -- not re-measured",
                         "kernel": "one solve of the batch = the k_phase<begin|prep|sweep|forward> launches of the phased driver (k_solve when the persistent kernel is selected); the backward-sweep kernel is 58 % of it (profiles/r01j_phased_solve_launches.json)",
                         "kernel_ms": kernel_ms,
                         "flop_per_launch": flop_launch,
                         "peak_source": "measured in this run by hsddp_fp64_peak_tflops: DFMA %.1f, DMMA %.1f TFLOP/s "
                                        "(MEASURED_PEAKS.json has no FP64 entry)" % (peak_dfma, peak_dmma),
                         "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                                 "bytes_per_launch": bytes_launch, "peak_source": hbm_src}},
            "latency": latency,
            "convergence": {k: float(v) for k, v in tot.items()},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_run(wl, pkg, w, args.cpu_seconds)
            line["cpu_baseline"] = {"value": r["value"], "unit": "solves/s", "cores": r["cores"], "kind": r["kind"],
                                    "sample": r["sample"], "model": r["model"]}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
