#!/usr/bin/env python3
"""bench.py -- batched QP-IK control cycles/s (FR3, batch 65536 per GPU), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

One "step" = one full control cycle for every robot of the batch:
    RobotData::updateState  +  RobotController::QPIKStep   (reference call sequence)
i.e. FK / Jacobians / CRBA / RNEA / M^-1 -> task error -> manipulability + self-collision rows ->
23-variable / 39-row QP solved with OSQP's algorithm -> qdot* (zeros on failure).

Printed JSON (one line, rank 0):
  value      whole-job cycles/s with inputs resident in HBM, device-timed (CUDA events on the launch
             stream, L2 flushed between steps, max over ranks)
  e2e        the same metric through the host-buffer C-ABI call (drc_host_cycle_qpik_step): pinned
             host inputs -> H2D -> kernels -> D2H inside the timed region
  roofline   dominant kernel (ADMM): algorithmic fp64 flops / measured duration vs the FP64 FMA peak
             measured on this device by drc_bench_fp64_peak (the path is fp64-ALU bound, not HBM or
             tensor bound: BASELINE.md section 4); the HBM view is reported next to it
  cpu_baseline  the oracle (CPU restatement, kind "port": the reference cannot be built here) timed
             on the box's host cores on a bounded sample
--impl reference times the oracle port on all host threads (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "batched QP-IK control cycles/sec (FR3, batch 65536)"  # metric name of the default workload; others append theirs
UNIT = "cycles/s"
LINK = "fr3_link8"
# Algorithmic fp64 flop model of the structured ADMM kernel for FR3 QPIK (NC=7, KU=2, ND=2; FMA = 2 flops),
# derived in DESIGN.md section "ADMM flop model": per iteration, per termination check, per (re)factorisation,
# and the one-off Ruiz scaling.
FLOPS_ITER, FLOPS_CHECK, FLOPS_FACTOR, FLOPS_SCALE = 867.0, 1450.0, 1900.0, 4300.0   # fallback (round-1 hand count, FR3 QPIK)
BYTES_PER_CYCLE = 320.0  # q, qdot, x_target(12), xdot_target in; qdot*, status, iters out


def flops_model(workload: str):
    """Algorithmic fp64 flop counts per robot from bench/flops_model.json (written by tools/count_flops.py: the oracle run with a
    flop-counting scalar on the benchmark inputs): ADMM per iteration / termination check / factorisation / equilibration, and the
    per-cycle front stages (kinematics, dynamics, manipulability, self-collision, QP build)."""
    try:
        m = json.loads((ROOT / "bench" / "flops_model.json").read_text())["workloads"].get(workload)
    except Exception:
        m = None
    if m is None and workload == "fr3_qpik":
        m = dict(admm=dict(iter=FLOPS_ITER, check=FLOPS_CHECK, factor=FLOPS_FACTOR, scale=FLOPS_SCALE), front=None, source="fallback hand count")
    return m


def ncu_traffic(workload: str, B: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_admm launch from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json, keyed by the source hash of the build it was taken from); None when the library has changed since."""
    try:
        from dyros_robot_controller_b200 import build as _b
        t = json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text())
        e = t.get(f"{workload}@{B}")
        return float(e["k_admm_dram_bytes"]) if (e and e.get("source_hash") == _b._source_hash()) else None
    except Exception:
        return None


def admm_flops(model, iters):
    """sum over robots of the ADMM flops for their iteration counts (rho updates: see flops_model.json 'refactor_rate')."""
    a = model["admm"]
    checks = np.ceil(iters / 25.0)
    refactors = 1.0 + a.get("refactor_rate", 0.01) * iters   # measured mean rho updates per iteration on the benchmark inputs
    return float(np.sum(a["scale"] + a["factor"] * refactors + a["iter"] * iters + a["check"] * checks))


# ---- workloads: "fr3_qpik" is BASELINE.json's metric (config 1 at the headline batch); the others are the sibling configs
# (3: fr3_qpid, 4: husky_*, 5: xls_*) for `--workload`, never the default line.
WORKLOADS = {
    "fr3_qpik": dict(robot="fr3", kind="ik", desc="FR3 updateState+QPIKStep (QP 23 vars / 39 rows, OSQP algorithm, self-collision + manipulability rows)"),
    "fr3_qpid": dict(robot="fr3", kind="id", desc="FR3 updateState+QPIDStep (QP 44 vars / 81 rows)"),
    "husky_qpik": dict(robot="husky_fr3", kind="ik", desc="Husky-FR3 (diff drive, synthesized URDF) whole-body updateState+QPIKStep (9 vars / 25 rows)"),
    "husky_qpid": dict(robot="husky_fr3", kind="id", desc="Husky-FR3 whole-body updateState+QPIDStep (18 vars / 39 rows)"),
    "xls_qpik": dict(robot="xls_fr3", kind="ik", desc="XLS-FR3 (mecanum, synthesized URDF) whole-body updateState+QPIKStep (11 vars / 27 rows)"),
    "xls_qpid": dict(robot="xls_fr3", kind="id", desc="XLS-FR3 whole-body updateState+QPIDStep (22 vars / 41 rows)"),
    "ur5e_clik_osf": dict(robot="ur5e", kind="taskspace", link="tool0",
                          desc="UR5e (synthesized URDF) updateState+CLIKStep+OSFStep, no QP (BASELINE config 2: run with --batch 4096)"),
}
MOMA_DESC = {
    "husky_fr3": dict(kin=dict(type="Differential", wheel_radius=0.1651, base_width=0.555), w=2,
                      joint_idx=dict(virtual_start=0, mobi_start=3, mani_start=5), actuator_idx=dict(mobi_start=0, mani_start=2)),
    "xls_fr3": dict(kin=dict(type="Mecanum", wheel_radius=0.120, roller_angles=[-np.pi / 4, np.pi / 4, np.pi / 4, -np.pi / 4],
                             base2wheel_positions=[(0.2225, 0.2045), (0.2225, -0.2045), (-0.2225, 0.2045), (-0.2225, -0.2045)],
                             base2wheel_angles=[0.0] * 4), w=4,
                    joint_idx=dict(virtual_start=0, mobi_start=3, mani_start=7), actuator_idx=dict(mobi_start=0, mani_start=4)),
}


def robot_paths(robot):
    import dyros_robot_controller_b200 as drc
    d = Path(drc.FR3_URDF).parents[1] / robot
    return str(d / f"{robot}.urdf"), str(d / f"{robot}.srdf")


def make_moma_workload(lo, hi, vl, w, B, seed):
    """base pose U([-2,2]^2 x [-pi,pi]), wheel angles U(-pi,pi), wheel speeds U(-2,2), arm as config 1."""
    rng = np.random.default_rng(seed)
    n = len(lo)
    al, ah, av = lo[3 + w:], hi[3 + w:], vl[3 + w:]
    q, qd = np.zeros((B, n)), np.zeros((B, n))
    q[:, 0:2] = rng.uniform(-2, 2, (B, 2)); q[:, 2] = rng.uniform(-np.pi, np.pi, B)
    q[:, 3:3 + w] = rng.uniform(-np.pi, np.pi, (B, w))
    q[:, 3 + w:] = al + (0.1 + 0.8 * rng.random((B, n - 3 - w))) * (ah - al)
    qd[:, 3:3 + w] = rng.uniform(-2, 2, (B, w))
    qd[:, 3 + w:] = rng.uniform(-0.5, 0.5, (B, n - 3 - w)) * av
    q_t = q.copy()
    q_t[:, 3 + w:] += 0.05 * rng.normal(size=(B, n - 3 - w))
    q_t[:, 0:2] += 0.05 * rng.normal(size=(B, 2))
    return q, qd, q_t, 0.05 * rng.normal(size=(B, 6))


def make_workload(model, B: int, seed: int):
    """BASELINE.md section 4 inputs (numpy default_rng(seed))."""
    rng = np.random.default_rng(seed)
    lo, hi, vl = model.q_lower, model.q_upper, model.v_limit
    q = lo + (0.1 + 0.8 * rng.random((B, model.dof))) * (hi - lo)
    qd = rng.uniform(-0.5, 0.5, (B, model.dof)) * vl
    dq = rng.normal(size=(B, model.dof)) * 0.05
    xdot_t = rng.normal(size=(B, 6)) * 0.05
    return q, qd, q + dq, xdot_t


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line): NVML every
    10 ms when pynvml is importable, else the nvidia-smi query every 200 ms."""

    _Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
          "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
          "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.sm, self.mx, self.power, self.reasons = [], [], [], set()
        self._stop_evt = threading.Event()
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(gpu_index))
            self._nvml = pynvml
        except Exception:
            self._nvml = None

    @staticmethod
    def _physical_index(i: int) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[i])
            except Exception:
                return i
        return i

    def _sample_nvml(self):
        n = self._nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
        self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self._h, n.NVML_CLOCK_SM)))
        try:
            self.power.append(n.nvmlDeviceGetPowerUsage(self._h) / 1000.0)
        except Exception:
            pass
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
            else n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        for name, bit in (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40),
                          ("sw_power_cap", 0x4)):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self._Q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if not out:
            return
        r = [c.strip() for c in out.split(",")]
        self.sm.append(float(r[0])); self.mx.append(float(r[1]))
        try:
            self.power.append(float(r[2]))
        except Exception:
            pass
        for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
            if val.lower().startswith("active"):
                self.reasons.add(name)

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop_evt.wait(0.01 if self._nvml is not None else 0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_mhz_min": min(self.sm) if self.sm else None,
                "sm_max_mhz": max(self.mx) if self.mx else None, "power_w_max": max(self.power) if self.power else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "source": "nvml" if self._nvml is not None else "nvidia-smi"}


def latency_b1(ctx, n_cpu: int = 3000, n_gpu: int = 1000):
    """BASELINE configs[0]: ONE FR3 updateState + QPIKStep control cycle (batch 1), the case the reference itself runs -- latency of
    the CPU port in reference-faithful mode on one thread (median and p99 after a warm-up) next to the latency of the same
    cycle through drc_host_cycle_qpik_step on the GPU (host buffers in, host buffers out, synchronous).  Different robot state per call."""
    from oracle.c_oracle import Oracle
    urdf, srdf = robot_paths("fr3")
    o = Oracle(urdf, srdf, threads=1)
    f = o.frame_id(LINK)

    class M:
        dof, q_lower, q_upper, v_limit = o.nv, o.model.q_lo, o.model.q_hi, o.model.v_lim
    n = max(n_cpu, n_gpu)
    q, qd, q_t, xdot_t = make_workload(M, n, 123)
    x_t = o.update_state(q_t, qd, f)["pose"]
    o.set_fresh_workspace(True)
    o.set_geom_params(gjk_tol=1e-6)
    for i in range(200):
        o.cycle(1, q[i:i + 1], qd[i:i + 1], x_t[i:i + 1], xdot_t[i:i + 1], f)
    tc = np.empty(n_cpu)
    for i in range(n_cpu):
        t0 = time.perf_counter()
        o.cycle(1, q[i:i + 1], qd[i:i + 1], x_t[i:i + 1], xdot_t[i:i + 1], f)
        tc[i] = time.perf_counter() - t0
    for i in range(50):
        ctx.cycle_qpik_step(q[i:i + 1], qd[i:i + 1], x_t[i:i + 1], xdot_t[i:i + 1], LINK)
    tg = np.empty(n_gpu)
    for i in range(n_gpu):
        t0 = time.perf_counter()
        ctx.cycle_qpik_step(q[i:i + 1], qd[i:i + 1], x_t[i:i + 1], xdot_t[i:i + 1], LINK)
        tg[i] = time.perf_counter() - t0
    us = lambda a, p_: float(np.percentile(a, p_) * 1e6)
    return {"workload": "FR3 updateState+QPIKStep, batch 1 (BASELINE configs[0]); per-call wall time incl. the Python / ctypes call",
            "cpu_us": {"median": us(tc, 50), "p99": us(tc, 99), "calls": n_cpu, "mode": "oracle port, one thread, reference-faithful"},
            "gpu_us": {"median": us(tg, 50), "p99": us(tg, 99), "calls": n_gpu, "path": "drc_host_cycle_qpik_step, pageable host buffers"}}


def oracle_cycles_per_s(B: int, threads: int, seed: int = 0, passes: int = 1, workload: str = "fr3_qpik", faithful: bool = False):
    """Time the CPU restatement (oracle port) of the workload on `threads` host threads.  faithful: the reference's per-cycle
    behaviour -- the whole workspace (solver data, scratch) heap-allocated and released every control cycle (QP_base.h:143-177,
    robot_data.cpp:542) and hpp-fcl's default GJK tolerance 1e-6 instead of the oracle's 1e-10."""
    from oracle.c_oracle import MomaOracle, Oracle
    wl = WORKLOADS[workload]
    urdf, srdf = robot_paths(wl["robot"])
    if wl["robot"] in ("fr3", "ur5e"):
        o = Oracle(urdf, srdf, threads=threads)
        f = o.frame_id(wl.get("link", LINK))

        class M:  # the oracle's own model view, same fields as engine.Model
            dof, q_lower, q_upper, v_limit = o.nv, o.model.q_lo, o.model.q_hi, o.model.v_lim
        q, qd, q_t, xdot_t = make_workload(M, B, seed)
        mode = 1 if wl["kind"] == "ik" else 3
        if wl["kind"] == "taskspace":
            class R:  # CLIKStep + OSFStep results; no QP iterations
                pass
            def run(*a):
                o.taskspace(0, *a, f)
                o.taskspace(1, *a, f)
                return dict(iters=np.zeros(len(a[0]), np.int32), status=np.ones(len(a[0]), np.int32))
        else:
            run = lambda *a: o.cycle(mode, *a, f)
    else:
        md = MOMA_DESC[wl["robot"]]
        o = MomaOracle(urdf, srdf, md["kin"], md["joint_idx"], md["actuator_idx"], threads=threads)
        f = o.frame_id(LINK)
        q, qd, q_t, xdot_t = make_moma_workload(o.model.q_lo, o.model.q_hi, o.model.v_lim, md["w"], B, seed)
        mode = 1 if wl["kind"] == "ik" else 3
        run = lambda *a: o.moma_cycle(mode, *a, f)
    x_t = o.update_state(q_t, qd, f)["pose"]
    if faithful:
        o.set_fresh_workspace(True)
        o.set_geom_params(gjk_tol=1e-6)
    run(q[:256], qd[:256], x_t[:256], xdot_t[:256])  # warm-up
    t0 = time.perf_counter()
    for _ in range(passes):
        r = run(q, qd, x_t, xdot_t)
    dt = time.perf_counter() - t0
    return B * passes / dt, dt / passes, r


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = args.batch   # same config as the product arm: one step = the whole batch (65536 robots take ~0.6 s on 16 threads)
    times = []
    for i in range(args.warmup + args.steps):
        cps, dt, _ = oracle_cycles_per_s(sample, threads, seed=i, workload=args.workload, faithful=True)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = sample / (ms * 1e-3)
    line = {"impl": "reference", "metric": METRIC if args.workload == "fr3_qpik" else f"batched control cycles/sec ({args.workload})", "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload]["desc"] + " -- CPU oracle port of the Pinocchio+OSQP path",
                       "batch_per_step": sample, "batch_per_gpu": sample, "global_batch": sample,
                       "mode": "reference-faithful: workspace allocated and released every cycle, GJK tolerance 1e-6 (hpp-fcl default)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{sample} cycles per step (the whole batch), OpenMP over all {threads} host threads, reference-faithful mode"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import dyros_robot_controller_b200 as drc

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    B = args.batch
    wl = WORKLOADS[args.workload]
    urdf, srdf = robot_paths(wl["robot"])
    model = drc.Model(urdf, srdf)
    moma = wl["robot"] not in ("fr3", "ur5e")
    link = wl.get("link", LINK)
    taskspace = wl["kind"] == "taskspace"
    if moma:
        md = MOMA_DESC[wl["robot"]]
        model.attach_mobile_base(md["kin"], md["joint_idx"], md["actuator_idx"])
    ctx = drc.Context(model, B, device=local)
    if os.environ.get("DRC_DEBUG_MAX_ITER"):  # experiment knob (tail analysis); never set for a reported number
        ctx.set_params(max_iter=int(os.environ["DRC_DEBUG_MAX_ITER"]))
    if os.environ.get("DRC_DEBUG_EPS"):       # experiment knob: force every robot to max_iter (lone-warp latency measurement)
        ctx.set_params(eps_abs=float(os.environ["DRC_DEBUG_EPS"]), eps_rel=float(os.environ["DRC_DEBUG_EPS"]))
    # each rank owns an independent shard of the batch (no exchange on the solve path)
    if moma:
        q, qd, q_t, xdot_t = make_moma_workload(model.q_lower, model.q_upper, model.v_limit, md["w"], B, seed=1000 * rank)
        ctx.moma_update_state(q_t, qd)
        x_t = ctx.moma_get_state(LINK, want=("pose",))["pose"]
        nout = model.actuated_dof
    else:
        # DRC_BENCH_SAME_SEED: experiment knob for the scaling analysis (every rank gets rank 0's batch: what is left of the per-rank
        # spread is the machine, not the data); never set for a reported number
        q, qd, q_t, xdot_t = make_workload(model, B, seed=0 if os.environ.get("DRC_BENCH_SAME_SEED") else 1000 * rank)
        ctx.update_state(q_t, qd)
        x_t = ctx.get_frame(link, want=("pose",))["pose"]
        nout = model.dof
    tq, tqd, txt, txd = (torch.from_numpy(a).to(dev) for a in (q, qd, x_t, xdot_t))
    out = torch.empty((B, nout), dtype=torch.float64, device=dev)
    out2 = torch.empty((B, nout), dtype=torch.float64, device=dev)
    st = torch.empty(B, dtype=torch.int32, device=dev)
    it = torch.empty(B, dtype=torch.int32, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def cycle(a_q, a_qd, a_xt, a_xd, o_out, o_st, o_it, o_out2=None):
        if moma:
            ctx.moma_cycle(wl["kind"], a_q, a_qd, a_xt, a_xd, LINK, out=o_out, out2=o_out2, status=o_st, iters=o_it)
        elif taskspace:   # config 2: updateState + CLIKStep (qdot*) + OSFStep (tau*), no QP -- one launch
            ctx.cycle_clik_osf_step(a_q, a_qd, a_xt, a_xd, link, out=o_out, out2=o_out2)
        elif wl["kind"] == "ik":
            ctx.cycle_qpik_step(a_q, a_qd, a_xt, a_xd, LINK, out=o_out, status=o_st, iters=o_it)
        else:
            ctx.cycle_qpid_step(a_q, a_qd, a_xt, a_xd, LINK, out=o_out, status=o_st, iters=o_it)

    # consecutive control ticks: between steps every robot's state advances by qdot * 1 ms (the reference's control period,
    # examples/robots/fr3/fr3.xml:4), so no two steps solve the same batch and the ADMM schedule hint (previous tick's
    # iteration counts) is exercised the way a control loop exercises it
    DT = 1e-3
    nticks = max(args.warmup, 3) + args.steps
    tq_k = [tq + (k * DT) * tqd for k in range(nticks)]

    def step(k):
        cycle(tq_k[k], tqd, txt, txd, out, st, it, out2)

    for k in range(max(args.warmup, 3)):
        step(k)
    torch.cuda.synchronize()
    ctx.enable_timing(not taskspace)
    sampler = ClockSampler(local)
    sampler.start()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = ctx.launch_count
    # the K timed steps are enqueued back to back, as a control pipeline would run them: each step is bracketed by its own pair
    # of CUDA events on the launch stream (the L2 flush sits between the pairs, outside every bracket), the host synchronises
    # once after the last step.  Steps cannot overlap (one stream orders them: a step's first kernel follows the previous step's
    # last), but the host is not stalled between them, so its ~30 driver calls per step hide behind the previous step's kernels.
    pairs = []
    for k in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step(max(args.warmup, 3) + k)
        e1.record()
        pairs.append((e0, e1))
    torch.cuda.synchronize()
    step_ms = [a.elapsed_time(b) for a, b in pairs]
    launches = ctx.launch_count - launches0
    # stage times (the library's own events, read back after every step): a short instrumented pass on the last ticks,
    # synchronised per step and therefore NOT part of the headline
    stage_ms = []
    for k in range(min(args.steps, 4)):
        flush.zero_()
        step(max(args.warmup, 3) + args.steps - 1 - k)
        torch.cuda.synchronize()
        stage_ms.append(ctx.last_timing() if not taskspace else dict(build_ms=0.0, collision_ms=0.0, admm_ms=0.0, total_ms=0.0))
    trace = ctx.last_trace() if not taskspace else []
    # the ADMM launch with nothing next to it: updateState, then QPIKStep / QPIDStep from the cached state (no priority pipeline, no
    # concurrent build / dynamics kernels; still in schedule order).  In the fused cycle above the main launch shares the GPU with the
    # priority launch, the dynamics-only kernel and the EPA-pending robots, so its in-situ duration (roofline.kernel_ms) is longer.
    alone = None
    if world == 1 and not taskspace and not moma:
        k_last = max(args.warmup, 3) + args.steps - 1
        q_np = q + (k_last * DT) * qd
        ts = []
        for rep in range(2):
            ctx.update_state(q_np, qd)
            ra = (ctx.qpik_step if wl["kind"] == "ik" else ctx.qpid_step)(x_t, xdot_t, LINK)
            ts.append(ctx.last_timing()["admm_ms"])
        alone = dict(kernel_ms=float(ts[-1]), iters=ra["iters"].astype(np.float64))
    if dist is not None:
        dist.barrier()
    clocks = sampler.stop()
    total_ms = float(np.sum(step_ms))
    if dist is not None:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * B / (ms_per_step * 1e-3)
    if taskspace:
        it.zero_(); st.fill_(1)
    iters_last, status_last = it.cpu().numpy().astype(np.float64), st.cpu().numpy()
    # the same ticks with the schedule hint off (identity robot order), reported next to the headline for transparency
    ctx.set_params(schedule_hint=0)
    nh = []
    for k in range(min(args.steps, 3)):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step(max(args.warmup, 3) + k)
        e1.record()
        e1.synchronize()
        nh.append(e0.elapsed_time(e1))
    ctx.set_params(schedule_hint=1)
    value_no_hint = world * B / (float(np.mean(nh)) * 1e-3)

    # ---- e2e through the host-buffer C-ABI call (pinned host memory in, host memory out)
    hq, hqd, hxt, hxd = (torch.from_numpy(a).pin_memory().numpy() for a in (q, qd, x_t, xdot_t))
    hq_k = [torch.from_numpy(q + (k * DT) * qd).pin_memory().numpy() for k in range(2 + args.steps)]
    hout = torch.empty((B, nout), dtype=torch.float64).pin_memory().numpy()
    hout2 = torch.empty((B, nout), dtype=torch.float64).pin_memory().numpy()
    hst = torch.empty(B, dtype=torch.int32).pin_memory().numpy()
    hit = torch.empty(B, dtype=torch.int32).pin_memory().numpy()
    ctx.enable_timing(False)
    for k in range(2):
        cycle(hq_k[k], hqd, hxt, hxd, hout, hst, hit, hout2)
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        cycle(hq_k[2 + k], hqd, hxt, hxd, hout, hst, hit, hout2)
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * B * args.steps / e2e_s
    h2d = B * (model.dof * 2 + 12 + 6) * 8
    d2h = B * (nout * 8 * (2 if ((moma and wl['kind'] == 'id') or taskspace) else 1) + (0 if taskspace else 8))

    # ---- per-rank evidence for the scaling curve: every rank's own step time, ADMM stage time, slowest robot
    per_rank = None
    if dist is not None:
        mine = torch.tensor([float(np.sum(step_ms)) / args.steps, float(np.min(step_ms)), float(np.median(step_ms)), float(np.max(step_ms)),
                             float(np.mean([s_["admm_ms"] for s_ in stage_ms])), float(np.mean([s_["collision_ms"] for s_ in stage_ms])),
                             float(np.mean([s_["build_ms"] for s_ in stage_ms])), float(iters_last.max()), float(iters_last.mean()),
                             float((iters_last >= 1000).sum()), e2e_s * 1e3 / args.steps], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        keys = ("ms_per_step", "step_ms_min", "step_ms_median", "step_ms_max", "admm_ms", "collision_ms", "build_ms", "max_admm_iters",
                "mean_admm_iters", "robots_ge_1000_iters", "e2e_ms_per_step")
        per_rank = [dict(zip(keys, [round(float(v), 4) for v in t.tolist()])) for t in allr]
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel (ADMM), algorithmic flops from the per-robot iteration counts
    iters, status = iters_last, status_last
    fm = flops_model(args.workload)
    flops = admm_flops(fm, iters) if (fm and not taskspace) else None
    admm_ms = float(np.mean([s["admm_ms"] for s in stage_ms]))
    build_ms = float(np.mean([s["build_ms"] for s in stage_ms]))
    col_ms = float(np.mean([s["collision_ms"] for s in stage_ms]))
    try:
        peak = drc.fp64_peak_tflops(local)
        peak_src = "measured in-run by drc_bench_fp64_peak (FP64 FMA, 8 chains/thread)"
    except Exception as e:  # pragma: no cover
        peak, peak_src = 37.0, f"fallback nominal B200 FP64 ({e})"
    # whole-body calls: the stage time above runs from the solver's dispatch to the end of the call and so includes the EPA pass, the
    # EPA-pending robots and the dynamics-only job (side streams, joined at the end); the main solver launch itself is bracketed by
    # the library's trace marks `admm_begin` / `admm` (last instrumented step)
    tr = dict(trace)
    if moma and tr.get("admm") is not None and tr.get("admm_begin") is not None and tr["admm"] > tr["admm_begin"]:
        admm_ms = float(tr["admm"] - tr["admm_begin"])
    achieved = flops / (admm_ms * 1e-3) / 1e12 if (flops is not None and admm_ms > 0) else None
    # the same launch under the round-1 hand-count model (867 / 1450 / 1900 / 4300 flop), for continuity with BENCH_r01
    r1_flops = float(np.sum(FLOPS_SCALE + FLOPS_FACTOR * (1.0 + np.floor(iters / 50.0) * 0.5) + FLOPS_ITER * iters + FLOPS_CHECK * np.ceil(iters / 25.0)))
    achieved_r1 = r1_flops / (admm_ms * 1e-3) / 1e12 if (args.workload == "fr3_qpik" and admm_ms > 0) else None
    # whole-step view: front stages (kinematics, dynamics, manipulability, self-collision, QP build) + ADMM over the step time
    step_flops = (flops + fm["front"]["total"] * B) if (flops is not None and fm.get("front")) else None
    step_achieved = step_flops / (ms_per_step * 1e-3) / 1e12 if step_flops is not None else None
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_achieved = BYTES_PER_CYCLE * B / (ms_per_step * 1e-3) / 1e9
    roofline = {"bound": "fp64", "kernel": "k_admm" + ("<QpCfg<7,2,2,0>>" if args.workload == "fr3_qpik" else ""), "achieved": achieved,
                "peak": peak, "unit": "TFLOP/s", "frac": (achieved / peak) if achieved is not None else None,
                # dram__bytes_read.sum + dram__bytes_write.sum of one k_admm launch at this batch from the ncu --set full capture
                # of the final build (profiles/r01_ncu_summary_v5_dynamics_split.md: 41.83 MB + 0.39 MB); algorithmic: 632 B record +
                # 64 B result per robot
                "traffic": ncu_traffic(args.workload, B),
                "traffic_algorithmic": (632.0 + 64.0) * B if args.workload == "fr3_qpik" else None,
                "peak_source": peak_src,
                "flops_model": (fm or {}).get("source"),
                "frac_round1_hand_model": (achieved_r1 / peak) if achieved_r1 is not None else None,
                "whole_step": {"achieved": step_achieved, "frac": (step_achieved / peak) if step_achieved is not None else None,
                               "flops_per_cycle": (step_flops / B) if step_flops is not None else None},
                "kernel_ms": admm_ms, "kernel_share_of_step": admm_ms / ms_per_step,
                # the same launch with nothing next to it (unfused QPIKStep from the cached state): duration, TFLOP/s, fraction of the peak
                "alone": ({"kernel_ms": alone["kernel_ms"], "achieved": admm_flops(fm, alone["iters"]) / (alone["kernel_ms"] * 1e-3) / 1e12,
                           "frac": admm_flops(fm, alone["iters"]) / (alone["kernel_ms"] * 1e-3) / 1e12 / peak}
                          if (alone and fm and alone["kernel_ms"] > 0) else None),
                "stage_ms": {"state_and_qp_build": build_ms, "self_collision": col_ms, "admm": float(np.mean([s_["admm_ms"] for s_ in stage_ms]))},
                # end time [ms since the call started] of every stage of the last instrumented step, main and priority pipeline
                "trace_ms": {k: round(v, 4) for k, v in trace},
                "hbm": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s"}}
    # ---- CPU baseline: the oracle port on this box's host cores, bounded sample
    cores = os.cpu_count() or 1
    sample = min(B, 32768)
    if world == 1:
        cpu_val, cpu_dt, _ = oracle_cycles_per_s(sample, cores, workload=args.workload)
        # the reference's control loop is single-threaded (one controller instance per control thread): one-thread rate on a smaller sample
        cpu1_val, cpu1_dt, _ = oracle_cycles_per_s(min(sample, 4096), 1, workload=args.workload)
        cpuf_val, cpuf_dt, _ = oracle_cycles_per_s(sample, cores, workload=args.workload, faithful=True)
    else:  # the CPU baseline is reported by the N = 1 run only
        cpu_val = cpu_dt = cpu1_val = cpu1_dt = cpuf_val = cpuf_dt = float("nan")
    line = {"metric": METRIC if args.workload == "fr3_qpik" else f"batched control cycles/sec ({args.workload})", "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            # this rank's timed steps one by one: the spread is the convergence tail (a step that holds an unpredicted max_iter robot)
            "step_ms": {"min": float(np.min(step_ms)), "median": float(np.median(step_ms)), "max": float(np.max(step_ms)),
                        "all": [round(float(x), 3) for x in step_ms]},
            "config": {"workload": wl["desc"], "batch_per_gpu": B, "global_batch": world * B,
                       "parallelism": f"batch shard x{world}, no collective on the solve path",
                       "l2": "256 MiB buffer zeroed between timed steps", "seed": "default_rng(1000*rank)",
                       "timing": "per-step CUDA event pairs on the launch stream, steps enqueued back to back, one host sync after the last",
                       "ticks": "consecutive control ticks: q advances by qdot*1ms between steps (no step repeats a batch)",
                       "admm_schedule": "robots ordered by the previous tick's iteration count (results unaffected)"},
            "value_no_schedule_hint": value_no_hint,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "peaks": {"fp64_tflops": peak, "fp64_source": peak_src, "hbm_gbs": hbm_peak, "hbm_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s"},
            "roofline": roofline,
            "cpu_baseline": None if world > 1 else {"value": cpu_val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{sample} cycles of the same workload, one pass, OpenMP over {cores} host threads "
                                       f"({cpu_dt:.2f} s)",
                             "reference_faithful": {"value": cpuf_val, "cores": cores,
                                                    "sample": f"{sample} cycles ({cpuf_dt:.2f} s); workspace allocated and released every cycle "
                                                              "(QP_base.h:143-177, robot_data.cpp:542), GJK tolerance 1e-6 (hpp-fcl default)"},
                             "single_thread": {"value": cpu1_val, "cores": 1,
                                               "sample": f"{min(sample, 4096)} cycles, one thread ({cpu1_dt:.2f} s); the reference's "
                                                         "control loop is single-threaded"}},
            "solved_fraction": float(np.mean(status == 1)), "mean_admm_iters": float(np.mean(iters))}
    if per_rank is not None:
        line["per_rank"] = per_rank
    if world == 1 and args.workload == "fr3_qpik":
        try:   # BASELINE configs[0] (batch 1): a latency pair, reported next to the throughput metric, never part of it
            ctx.enable_timing(False)
            line["latency_b1"] = latency_b1(ctx)
        except Exception as e:
            line["latency_b1"] = {"error": f"{type(e).__name__}: {e}"}
    if world == 1 and args.workload == "fr3_qpik" and not args.no_siblings:
        del ctx, flush
        torch.cuda.empty_cache()
        sib = {}
        for name, Bs in SIBLINGS:
            try:
                sib[name] = measure_sibling(name, Bs, local, peak)
            except Exception as e:  # a sibling must never cost the headline line
                sib[name] = {"error": f"{type(e).__name__}: {e}"}
        line["siblings"] = sib
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


# BASELINE.json configs 2-5 at their stated batch sizes (config 5: the full 1 M batch on one GPU), measured after the headline so
# that they are driver-run numbers too.  Shorter runs than the headline: 3 warm-up + 5 timed steps each.
SIBLINGS = [("ur5e_clik_osf", 4096), ("fr3_qpid", 65536), ("husky_qpik", 262144), ("husky_qpid", 262144),
            ("xls_qpik", 1048576), ("xls_qpid", 1048576)]


def measure_sibling(name: str, B: int, local: int, peak: float, steps: int = 5, warmup: int = 3):
    """one sibling configuration on this GPU: device-timed value, e2e through the host-buffer C-ABI, CPU baseline on a bounded sample,
    ADMM roofline from the flop model.  Same measurement rules as the headline (CUDA events on the launch stream, L2 flushed
    between steps, consecutive control ticks)."""
    import torch
    import dyros_robot_controller_b200 as drc
    wl = WORKLOADS[name]
    urdf, srdf = robot_paths(wl["robot"])
    model = drc.Model(urdf, srdf)
    moma = wl["robot"] not in ("fr3", "ur5e")
    link = wl.get("link", LINK)
    taskspace = wl["kind"] == "taskspace"
    dev = torch.device("cuda", local)
    if moma:
        md = MOMA_DESC[wl["robot"]]
        model.attach_mobile_base(md["kin"], md["joint_idx"], md["actuator_idx"])
    ctx = drc.Context(model, B, device=local)
    if moma:
        q, qd, q_t, xdot_t = make_moma_workload(model.q_lower, model.q_upper, model.v_limit, md["w"], B, seed=0)
        ctx.moma_update_state(q_t, qd)
        x_t = ctx.moma_get_state(LINK, want=("pose",))["pose"]
        nout = model.actuated_dof
    else:
        q, qd, q_t, xdot_t = make_workload(model, B, seed=0)
        ctx.update_state(q_t, qd)
        x_t = ctx.get_frame(link, want=("pose",))["pose"]
        nout = model.dof
    tq, tqd, txt, txd = (torch.from_numpy(a).to(dev) for a in (q, qd, x_t, xdot_t))
    out, out2 = (torch.empty((B, nout), dtype=torch.float64, device=dev) for _ in range(2))
    st, it = (torch.empty(B, dtype=torch.int32, device=dev) for _ in range(2))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def cycle(a_q, a_qd, a_xt, a_xd, o_out, o_st, o_it, o_out2):
        if moma:
            ctx.moma_cycle(wl["kind"], a_q, a_qd, a_xt, a_xd, LINK, out=o_out, out2=o_out2, status=o_st, iters=o_it)
        elif taskspace:
            ctx.cycle_clik_osf_step(a_q, a_qd, a_xt, a_xd, link, out=o_out, out2=o_out2)
        elif wl["kind"] == "ik":
            ctx.cycle_qpik_step(a_q, a_qd, a_xt, a_xd, LINK, out=o_out, status=o_st, iters=o_it)
        else:
            ctx.cycle_qpid_step(a_q, a_qd, a_xt, a_xd, LINK, out=o_out, status=o_st, iters=o_it)

    DT = 1e-3
    for k in range(warmup):
        cycle(tq + (k * DT) * tqd, tqd, txt, txd, out, st, it, out2)
    torch.cuda.synchronize()
    ctx.enable_timing(not taskspace)
    l0 = ctx.launch_count
    pairs = []
    for k in range(steps):
        qk = tq + ((warmup + k) * DT) * tqd
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        cycle(qk, tqd, txt, txd, out, st, it, out2)
        e1.record()
        pairs.append((e0, e1))
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in pairs]))
    launches = ctx.launch_count - l0
    stage = ctx.last_timing() if not taskspace else dict(build_ms=0.0, collision_ms=0.0, admm_ms=0.0, total_ms=0.0)
    try:
        trace = dict(ctx.last_trace()) if not taskspace else {}
    except Exception:
        trace = {}
    if taskspace:
        it.zero_(); st.fill_(1)
    iters, status = it.cpu().numpy().astype(np.float64), st.cpu().numpy()
    # e2e: pinned host buffers through the drc_host_* entry points
    ctx.enable_timing(False)
    hq, hqd, hxt, hxd = (torch.from_numpy(a).pin_memory().numpy() for a in (q, qd, x_t, xdot_t))
    hout, hout2 = (torch.empty((B, nout), dtype=torch.float64).pin_memory().numpy() for _ in range(2))
    hst, hit = (torch.empty(B, dtype=torch.int32).pin_memory().numpy() for _ in range(2))
    cycle(hq, hqd, hxt, hxd, hout, hst, hit, hout2)
    n_e2e = 3
    t0 = time.perf_counter()
    for k in range(n_e2e):
        cycle(hq, hqd, hxt, hxd, hout, hst, hit, hout2)
    e2e = B * n_e2e / (time.perf_counter() - t0)
    h2d = B * (model.dof * 2 + 12 + 6) * 8
    d2h = B * (nout * 8 * (2 if ((moma and wl["kind"] == "id") or taskspace) else 1) + (0 if taskspace else 8))
    # CPU baseline on a bounded sample of the same workload
    cores = os.cpu_count() or 1
    sample = min(B, 8192)
    cpu_val, cpu_dt, _ = oracle_cycles_per_s(sample, cores, workload=name)
    fm = flops_model(name)
    admm_ms = float(stage["admm_ms"])     # ev[2] -> ev[3]: from the solver's dispatch to the end of the call (joins the EPA pass, the dynamics job)
    # the main solver launch itself, in situ: the library's trace marks either side of it (last timed step).  The whole-body calls run
    # the EPA pass, the EPA-pending robots and the dynamics-only job next to / behind it, which the stage time above includes.
    kernel_ms = admm_ms
    if trace.get("admm") is not None and trace.get("admm_begin") is not None and trace["admm"] > trace["admm_begin"]:
        kernel_ms = float(trace["admm"] - trace["admm_begin"])
    fl = admm_flops(fm, iters) if (fm and not taskspace) else None
    ach = fl / (kernel_ms * 1e-3) / 1e12 if (fl and kernel_ms > 0) else None
    if taskspace and fm and fm.get("front"):   # no QP: the whole step is the front stage
        ach = fm["front"]["total"] * B / (ms * 1e-3) / 1e12
    del ctx, flush
    torch.cuda.empty_cache()
    return {"workload": wl["desc"], "batch": B, "value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "warmup": warmup,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "kernel": "k_admm" if not taskspace else "k_robot_job", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                         "frac": (ach / peak) if ach is not None else None, "kernel_ms": kernel_ms if not taskspace else ms,
                         "kernel_share_of_step": (kernel_ms / ms) if not taskspace else 1.0,
                         "stage_ms": {"state_and_qp_build": float(stage["build_ms"]), "self_collision": float(stage["collision_ms"]), "admm": admm_ms},
                         "trace_ms": {k: round(float(v), 4) for k, v in trace.items()},
                         "flops_model": (fm or {}).get("source")},
            "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{sample} cycles of the same workload, one pass, OpenMP over {cores} host threads ({cpu_dt:.2f} s)"},
            "solved_fraction": float(np.mean(status == 1)), "mean_admm_iters": float(np.mean(iters))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=65536, help="robots per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="fr3_qpik", choices=sorted(WORKLOADS), help="default = BASELINE.json's metric")
    ap.add_argument("--no-siblings", action="store_true", help="skip the sibling configurations (BASELINE configs 2-5) after the headline")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
