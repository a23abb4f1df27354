#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: FR3 RNEA+FD evals/sec (fp64) on N B200s, with roofline fractions.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...     # the reference arm: the CPU oracle on all host threads

One step = one batched RNEA pass (BASELINE.json configs[1]) + one batched forward-dynamics pass (configs[2]) over
2^24 synthetic FR3 states per GPU, device-resident SoA, sampled on the device by the counter-based generator of
SURVEY.md 8d.  Each GPU owns its own 2^24 states (weak scaling, no data-path collective).  Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "FR3 RNEA+FD evals/sec (fp64)"
UNIT = "evals/s"
STATES_PER_GPU = 1 << 24
# algorithmic work per evaluation (SURVEY.md 8d / BASELINE.md section 2), N = 7
RNEA_FLOPS, FD_FLOPS, BYTES_PER_EVAL = 2026.0, 4180.0, 224.0
SEED_RNEA, SEED_FD = 0x5EED0001, 0x5EED0002
# FP64-pipe SASS instructions of the fr3-specialised kernels (tools/sass_count.sh; tests/test_host.py checks the static
# totals against the built library).  The kernels are straight-line except for the out-of-range fallback of the batched
# sin/cos (rb_sincos_batched: six per-joint CUDA sincos() calls, 21 FP64 instructions each, taken only when some
# |q_i| >= 1e5): executed per state = static - fallback.  ncu's smsp__sass_thread_inst_executed_op_d{fma,add,mul}
# counters in profiles/ agree.
FP64_STATIC = {"rb_rnea_kernel": 804, "rb_fd_kernel": 1299}
FP64_FALLBACK = {"rb_rnea_kernel": 126, "rb_fd_kernel": 126}
FP64_INSTR = {k: FP64_STATIC[k] - FP64_FALLBACK[k] for k in FP64_STATIC}
FP64_LANES_PER_SM = 64


def profiled_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed ncu --set full capture."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    names = sorted(os.listdir(pdir), key=lambda nm: ("final" in nm, nm)) if os.path.isdir(pdir) else []
    for name in names:                                   # the last match wins: the capture named *final* if there is one
        if not name.endswith("ncu_full_summary.json"):
            continue
        with open(os.path.join(pdir, name)) as f:
            for row in json.load(f):
                if kernel in row.get("kernel", "") and "dram__bytes_read.sum" in row:
                    def gb(v):
                        num, unit = v.split()[:2]
                        return float(num) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[unit]
                    best = {"bytes": gb(row["dram__bytes_read.sum"]) + gb(row["dram__bytes_write.sum"]), "source": "profiles/" + name}
    return best


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock, power and throttle reasons of one GPU DURING the timed region (B200_PROFILING.md recipe).
    NVML through pynvml (a query takes ~0.1 ms, so even a 20 ms timed region gets several samples); falls back to
    polling nvidia-smi when pynvml is unavailable."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and all(x.strip().isdigit() for x in visible.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self._reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def run(self):
        if self.nvml is not None:
            nv = self.nvml
            while not self._stop_evt.is_set():
                try:
                    self.rows.append((float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)),
                                      nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0,
                                      int(self._reasons(self.handle))))
                except Exception:
                    pass
                self._stop_evt.wait(0.002)
            return
        fields = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                  "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={fields}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout
                p = [x.strip() for x in out.strip().split(",")]
                if len(p) == 7:
                    mask = sum(bit for bit, k in zip((0x8, 0x40, 0x20, 0x4), range(3, 7)) if p[k].lower().startswith("active"))
                    self.max_mhz = float(p[1])
                    self.rows.append((float(p[0]), float(p[2]), mask))
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(r[0] for r in self.rows)
        mask = 0
        for r in self.rows:
            mask |= r[2]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": getattr(self, "max_mhz", None),
                "reasons": [name for bit, name in self.REASONS.items() if mask & bit],
                "power_w_max": max(r[1] for r in self.rows), "samples": len(self.rows),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ----------------------------------------------------------------------------- CPU arm (oracle; test infrastructure)
def cpu_arm(sample_states, steps, warmup):
    """Times oracle/rb_oracle.c (reference-shaped C restatement; the Rust crate cannot be built in this image),
    OpenMP static chunks over all host threads: RNEA + FD over `sample_states` states per step."""
    from oracle.rb_oracle import Oracle
    orc = Oracle.from_urdf(os.path.join(ROOT, "assets", "fr3.urdf"), fast=True)   # rebuilt -march=native on this host
    m = orc.model
    # torchrun exports OMP_NUM_THREADS=1: size the team from the CPUs this process may run on instead
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if sample_states is None:       # calibrate to ~4 s of wall time per step
        S0 = 1 << 14
        q = orc.fill(SEED_RNEA, 0, m.lower, m.upper, 0, S0); dq = orc.fill(SEED_RNEA, 1, -m.velocity, m.velocity, 0, S0)
        t0 = time.perf_counter()
        orc.rnea_batch(q, dq, q, threads=threads); orc.forward_dynamics_batch(q, dq, q, threads=threads)
        per = (time.perf_counter() - t0) / S0
        sample_states = int(min(1 << 22, max(1 << 14, 4.0 / per)))
    S = sample_states
    q = orc.fill(SEED_RNEA, 0, m.lower, m.upper, 0, S)
    dq = orc.fill(SEED_RNEA, 1, -m.velocity, m.velocity, 0, S)
    ddq = orc.fill(SEED_RNEA, 2, -10.0, 10.0, 0, S)
    tau = orc.fill(SEED_FD, 3, -m.effort, m.effort, 0, S)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        orc.rnea_batch(q, dq, ddq, threads=threads)
        orc.forward_dynamics_batch(q, dq, tau, threads=threads)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    # BASELINE.json configs[0]: one state, one core (the reference's bench_rnea input, multibody.rs:204-209)
    z = np.zeros((7, 1)); reps = 20000
    t0 = time.perf_counter()
    for _ in range(3):
        orc.rnea_batch(np.repeat(z, reps, 1), np.repeat(z, reps, 1), np.repeat(z, reps, 1), threads=1)
    single_ns = (time.perf_counter() - t0) / (3 * reps) * 1e9
    return {"value": 2.0 * S / sec, "unit": UNIT, "cores": threads, "kind": "port", "single_state_rnea_ns_one_core": single_ns,
            "sample": f"{S} FR3 states per step (RNEA + FD each), {steps} timed steps, OpenMP static over {threads} threads, "
                      f"oracle/rb_oracle.c -O3 -march=native"}, sec, S


def run_reference(args, rank):
    if rank != 0:
        return
    cb, sec, S = cpu_arm(None, max(1, args.steps), max(1, args.warmup))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "fr3_rnea+fd", "states_per_step": S, "layout": "soa",
                       "note": "CPU port of the reference path (oracle/); the Rust crate cannot be built here"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import rigidbody_rs_b200 as rb
    from rigidbody_rs_b200.shard import max_over_ranks, sum_over_ranks

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    mb = rb.Multibody.from_urdf(os.path.join(ROOT, "assets", "fr3.urdf"), device=local_rank)
    n, B = mb.n, args.states
    lim = mb.limits()
    first = rank * B                                   # each rank samples its own slice of the global index space
    q = torch.empty((n, B), dtype=torch.float64, device=dev)
    dq, ddq, tau_in = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
    tau, qdd = torch.empty_like(q), torch.empty_like(q)
    mb.fill(q, SEED_RNEA, 0, lim["lower"], lim["upper"], first)
    mb.fill(dq, SEED_RNEA, 1, -lim["velocity"], lim["velocity"], first)
    mb.fill(ddq, SEED_RNEA, 2, -10.0, 10.0, first)
    mb.fill(tau_in, SEED_FD, 3, -lim["effort"], lim["effort"], first)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        mb.rnea(q, dq, ddq, out=tau)
        mb.forward_dynamics(q, dq, tau_in, out=qdd)

    for _ in range(args.warmup):
        step()
    K = args.steps
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    sampler = ClockSampler(local_rank); sampler.start()
    barrier()
    launches0 = mb.launch_count
    t_start = torch.cuda.Event(enable_timing=True); t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for k in range(K):
        ev[k][0].record()
        mb.rnea(q, dq, ddq, out=tau)
        ev[k][1].record()
        mb.forward_dynamics(q, dq, tau_in, out=qdd)
        ev[k][2].record()
    t_end.record()
    barrier()
    clocks = sampler.stop()
    launches = mb.launch_count - launches0
    mb.sync()
    total_ms = t_start.elapsed_time(t_end)
    # FP64 roofline denominator: DFMA probe on the same GPU in the same run, run for about as long as the timed
    # region (a burst of a few ms keeps 1965 MHz; seconds of FP64 work pull ~1 kW and settle near 1.7 GHz under
    # sw_power_cap), so burst numbers are divided by a burst peak and sustained numbers by a sustained peak
    probe_ms = int(min(2000.0, max(100.0, total_ms)))
    fp64_peak = mb.fp64_peak_tflops(probe_ms)
    ms_step = max_over_ranks(total_ms / K, dev)
    ms_rnea = max_over_ranks(sum(e[0].elapsed_time(e[1]) for e in ev) / K, dev)
    ms_fd = max_over_ranks(sum(e[1].elapsed_time(e[2]) for e in ev) / K, dev)
    units = sum_over_ranks(2.0 * B, dev)               # evaluations all ranks processed per step
    value = units / (ms_step * 1e-3)

    # ---- end to end through the public API with HOST buffers (pinned), H2D + D2H inside the timed region
    Be = args.e2e_states
    hq, hdq, hddq, htau_in = (rb.host_empty((n, Be)) for _ in range(4))
    htau, hqdd = rb.host_empty((n, Be)), rb.host_empty((n, Be))
    for h, d in ((hq, q), (hdq, dq), (hddq, ddq), (htau_in, tau_in)):
        h[...] = d[:, :Be].cpu().numpy()
    e2e_steps = max(1, min(K, args.e2e_steps))
    mb.rnea(hq, hdq, hddq, out=htau); mb.forward_dynamics(hq, hdq, htau_in, out=hqdd)      # warm-up (allocates staging)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        mb.rnea(hq, hdq, hddq, out=htau)
        mb.forward_dynamics(hq, hdq, htau_in, out=hqdd)
    torch.cuda.synchronize()
    sep_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / e2e_steps, dev)
    # the same step through the one-call entry point (multibody_rnea_fd_batch): identical kernels and results, but q and
    # dq cross PCIe once instead of twice (4 input arrays instead of 6) -- the bus is what bounds this number
    hout = rb.host_empty((2 * n, Be))
    mb.rnea_fd(hq, hdq, hddq, htau_in, out=hout)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        mb.rnea_fd(hq, hdq, hddq, htau_in, out=hout)
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / e2e_steps, dev)
    e2e_units = sum_over_ranks(2.0 * Be, dev)
    # the host results equal the device-resident ones bit for bit (same kernels, same inputs)
    same = bool(np.array_equal(htau[:, :4096], tau[:, :4096].cpu().numpy())) and bool(np.array_equal(hout[:n], htau)) \
        and bool(np.array_equal(hout[n:], hqdd))

    # optional assembly of the sharded result on every rank: ONE NCCL all-gather of tau, outside the timed
    # region and reported separately -- the compute path itself exchanges nothing
    gather = None
    if world > 1:
        Bg = min(B, 1 << 22)                              # 235 MB per rank: enough to see the NVLink rate
        part = tau[:, :Bg].contiguous()
        full = torch.empty((world, n, Bg), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(full, part)           # warm-up (communicator set-up)
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        g0.record()
        dist.all_gather_into_tensor(full, part)
        g1.record()
        torch.cuda.synchronize()
        gms = max_over_ranks(g0.elapsed_time(g1), dev)
        ok_g = bool(torch.equal(full[rank], part))
        gather = {"collective": "ncclAllGather(tau)", "ms": gms, "bytes_per_rank": int(part.numel() * 8),
                  "recv_GB_per_s_per_rank": (world - 1) * part.numel() * 8 / (gms * 1e-3) / 1e9, "own_slice_intact": ok_g}

    hbm_peak, hbm_src = measured_peaks()
    rnea_s, fd_s = ms_rnea * 1e-3, ms_fd * 1e-3

    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6

    def roof(kernel, flops, sec):
        tf = flops * B / sec / 1e12
        gbs = BYTES_PER_EVAL * B / sec / 1e9
        tr = profiled_traffic(kernel) if B == STATES_PER_GPU else None
        pipe = FP64_INSTR[kernel] * B / sec / (FP64_LANES_PER_SM * sm_count * sm_hz) if mb.kernel_variant == "fr3-specialised" else None
        return {"kernel": kernel, "bound": "fp64", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": tf / fp64_peak, "traffic": tr["bytes"] if tr else None, "traffic_source": tr["source"] if tr else None,
                "fp64_pipe_util": pipe, "fp64_instr_per_eval": FP64_INSTR[kernel],
                "evals_per_s": B / sec, "ms": sec * 1e3,
                "peak_source": f"DFMA probe (multibody_gpu_measure_fp64_peak) on this GPU in this run, {probe_ms} ms; nominal 37.2",
                "flops_per_eval": flops, "bytes_per_eval": BYTES_PER_EVAL,
                "hbm": {"achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "peak_source": hbm_src}}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "fr3_rnea+fd_16M", "states_per_gpu": B, "n_joints": n, "layout": "soa",
                   "kernel_variant": mb.kernel_variant, "cache": "inputs larger than L2 (2.8 GB read per kernel)",
                   "step": "1 RNEA launch + 1 forward-dynamics launch over all states", "parallelism": f"dp{world}"},
        "roofline": roof("rb_fd_kernel", FD_FLOPS, fd_s),
        "roofline_rnea": roof("rb_rnea_kernel", RNEA_FLOPS, rnea_s),
        "e2e": {"value": e2e_units / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 4 * n * Be * 8,
                "d2h_bytes_per_step": 2 * n * Be * 8, "states_per_gpu": Be, "steps": e2e_steps, "ms_per_step": e2e_ms,
                "api": "Multibody.rnea_fd on pinned numpy arrays -> multibody_rnea_fd_batch(RB_MEM_HOST): q, dq, ddq, tau_in up, tau and qdd down",
                "matches_device_path": same,
                "separate_calls": {"value": e2e_units / (sep_ms * 1e-3), "ms_per_step": sep_ms, "h2d_bytes_per_step": 6 * n * Be * 8,
                                   "api": "Multibody.rnea + Multibody.forward_dynamics -> multibody_{rnea,forward_dynamics}_batch(RB_MEM_HOST)"}},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if gather is not None:
        line["gather"] = gather
    if rank == 0 and world == 1 and not args.no_cpu:
        cb, _, _ = cpu_arm(None, 2, 1)
        line["cpu_baseline"] = cb
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line), flush=True)


def run_other(args, rank, local_rank, world):
    """BASELINE.json configs[3] (FR3 MPC rollout, 65 536 trajectories x 64 steps, strong scaling: the trajectories
    are split across ranks) and configs[4] (32-joint chain RNEA + FD, weak scaling).  Same JSON shape, kernel-only."""
    import torch
    import torch.distributed as dist
    import rigidbody_rs_b200 as rb
    from rigidbody_rs_b200.shard import max_over_ranks, shard_bounds, sum_over_ranks

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    rollout = args.workload == "rollout"
    mb = rb.Multibody.from_urdf(os.path.join(ROOT, "assets", "fr3.urdf" if rollout else "chain32.urdf"), device=local_rank)
    n, lim = mb.n, mb.limits()
    H = 64
    if rollout:
        lo, hi = shard_bounds(65536, world, rank)
        B, first, seed = hi - lo, lo, 0x5EED0003
    else:
        B = args.states if args.states != STATES_PER_GPU else 1 << 22
        first, seed = rank * B, 0x5EED0005
    q = torch.empty((n, B), dtype=torch.float64, device=dev)
    dq, x3, out, out2 = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
    mb.fill(q, seed, 0, lim["lower"], lim["upper"], first)
    mb.fill(dq, seed, 1, -lim["velocity"], lim["velocity"], first)
    if rollout:
        tau = torch.empty((H, n, B), dtype=torch.float64, device=dev)
        for t in range(H):
            mb.fill(tau[t], seed, 4 + t % 32, -lim["effort"], lim["effort"], t * 65536 + first)
        step = lambda: mb.rollout(q, dq, tau, 1e-3)
        units_rank, flops, label = float(B * H), 4208.0, "rb_rollout_kernel"
    else:
        mb.fill(x3, seed, 2, -10.0, 10.0, first)
        tau_in = torch.empty_like(q)
        mb.fill(tau_in, seed, 3, -lim["effort"], lim["effort"], first)
        def step():
            mb.rnea(q, dq, x3, out=out)
            mb.forward_dynamics(q, dq, tau_in, out=out2)
        units_rank, flops, label = 2.0 * B, (9476.0 + 53800.0) / 2, "rb_long_rnea_kernel + rbh_fd_kernel"
    for _ in range(args.warmup):
        step()
    fp64_peak = mb.fp64_peak_tflops(100)
    sampler = ClockSampler(local_rank); sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = mb.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = max_over_ranks(e0.elapsed_time(e1) / args.steps, dev)
    units = sum_over_ranks(units_rank, dev)
    tf = flops * units / world / (ms * 1e-3) / 1e12
    line = {"metric": "FR3 MPC rollout FD steps/sec (fp64)" if rollout else "chain32 RNEA+FD evals/sec (fp64)",
            "value": units / (ms * 1e-3), "unit": "steps/s" if rollout else UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if rollout else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "fr3_rollout_65536x64" if rollout else "chain32_rnea+fd", "states_per_gpu": B, "n_joints": n,
                       "horizon": H if rollout else None, "dt": 1e-3 if rollout else None, "kernel_variant": mb.kernel_variant,
                       "parallelism": f"dp{world}"},
            "roofline": {"kernel": label, "bound": "fp64", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak,
                         "traffic": None, "flops_per_unit": flops},
            "gpu_launches": int(mb.launch_count - l0), "clocks": clocks}
    if rank == 0:
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--states", type=int, default=STATES_PER_GPU, help="states per GPU per step")
    ap.add_argument("--e2e-states", type=int, default=STATES_PER_GPU, help="states per GPU per end-to-end step")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="fr3", choices=["fr3", "rollout", "chain32"],
                    help="fr3 = the headline RNEA+FD line (BASELINE.json configs[1]+[2]); rollout = configs[3]; chain32 = configs[4]")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU arm)")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        (run_ours if args.workload == "fr3" else run_other)(args, rank, local_rank, world)
    finally:
        if world > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
