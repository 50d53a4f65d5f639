#!/usr/bin/env python
"""bench.py -- particle-steps/s of the Vlasov-Poisson spline-PIC step on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...   # CPU arm: C restatement of the reference algorithm

Workload (config.workload): bump-on-tail, BASELINE.json configs[1]: 1e8 particles TOTAL (strong
scaling over N GPUs), periodic cubic B-splines (degree 3 = order 4), n_h = 16, L = 2 pi / 0.3,
dt = 0.1 (scripts/bump_on_tail.jl:14-32 scaled to 1e8 particles), synthetic Philox load.

A "step" = one Strang step: E gather -> kick -> drift -> charge deposition -> all-reduce -> Poisson solve.
`value`  : particles resident in HBM, K steps timed with CUDA events on the library's stream.
`e2e`    : same K steps through the host API with HOST (pinned) particle arrays: upload of x,v,w,
           K steps with a [W,K,M] diagnostics read-back, download of x,v -- all inside the timed region.
`roofline`: the fused push+deposit kernel, 32 algorithmic bytes per particle per launch (read x,v; write x,v;
           the uniform particle weight of this workload is a kernel parameter -- 40 B with `--general-weights`,
           which streams the weight array as for per-particle weights), timed per
           launch with CUDA event brackets over a second run of the same K steps (the brackets defeat
           the programmatic-dependent-launch overlap, so they stay out of the `value` region).
`deposit`: the deposit-only pass (projection! alone) with the uniform weight and with the weight array streamed.
`secondary`: LB / CLB right-hand sides and the CLB RK438 step at the same particle count (configs[2], [3]) and, on
           one GPU, the fused step on meshes of 32 ... 1024 cells (configs[4]).
`cpu_baseline`: the C restatement of the reference algorithm on the host cores (N = 1 only), on the SAME 1e8-particle
           arrays as the GPU arm's workload (a bounded number of steps), as is `--impl reference`.
`parity`   : a 40 000-particle sharded run on the same ranks checked against the CPU oracle (checker only) and the
           field coefficients compared bitwise across ranks -- the multi-GPU correctness record of the driver runs.
`value`    : median over `repeats` timed regions of K steps each (min / max alongside); before every region the ranks
           are aligned on the device by one untimed fused step (its in-kernel exchange waits for the slowest rank).
`ceilings` : tools/microbench/peaks (fp64 FMA rate, shared-memory wavefront rate, copy bandwidth) run on the same GPU.
`step_frac_of_in_run_copy_rate`: the step against that copy kernel -- same box, same power / clock state (`roofline.frac`
           and `step_hbm_frac` use the driver's cool-GPU peak of MEASURED_PEAKS.json, as the contract asks).
`secondary.mesh_sweep_n_basis[*].deposit_layout`: which deposition layout the planner chose for that mesh.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_TOTAL = 100_000_000
N_BASIS, ORDER, DT = 16, 4, 0.1
EPS, KAPPA, ALPHA, SIGMA, V0 = 0.03, 0.3, 0.1, 0.5, 4.5
L_DOMAIN = 2 * math.pi / KAPPA
# Algorithmic bytes per particle-step of the fused pass: read x,v (16 B) + write x,v (16 B),
# plus 8 B for the weight when weights differ per particle.  Every sampler of the reference (and this
# workload) gives all particles ONE weight (w = L/N), which the library detects and then passes w0 as a
# kernel parameter instead of streaming the array -- so the honest figure for this workload is 32 B.
ALG_BYTES_UNIFORM_W, ALG_BYTES_GENERAL_W = 32, 40
SEED = 20240601


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.sm_max = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            pass
    return local_rank


# ------------------------------------------------------------------ CPU arm --
def native_oracle():
    """-march=native build of the oracle for the timed CPU baseline (built on the machine that runs it)."""
    from oracle import vm_oracle as orc
    out = Path(tempfile.gettempdir()) / f"libvm_oracle_native_{os.getpid()}.so"
    try:
        orc.build(native_out=str(out))
        return orc, orc.lib(str(out))
    except Exception:
        return orc, orc.lib()


def _sample_chunk(args):
    lo, hi, seed = args
    rng = np.random.default_rng([seed, lo])
    m = hi - lo
    u = rng.uniform(size=m)
    x = u * L_DOMAIN
    for _ in range(8):          # Newton on the CDF x - (eps/kappa) sin(kappa x) = u L: eps = 0.03, quadratic convergence
        x -= (x - (EPS / KAPPA) * np.sin(KAPPA * x) - u * L_DOMAIN) / (1 - EPS * np.cos(KAPPA * x))
    v = rng.standard_normal(m)
    tail = rng.uniform(size=m) > 1 - ALPHA
    v[tail] = v[tail] * SIGMA + V0
    return lo, x, v


def sample_particles(n, seed=SEED):
    """Bump-on-tail load for the CPU arm (numpy, chunks on all host threads; same distribution as the device fill)."""
    from concurrent.futures import ThreadPoolExecutor
    x, v = np.empty(n), np.empty(n)
    step = 2_000_000
    with ThreadPoolExecutor(max_workers=host_threads()) as ex:
        for lo, xs, vs in ex.map(_sample_chunk, [(lo, min(n, lo + step), seed) for lo in range(0, n, step)]):
            x[lo:lo + xs.size], v[lo:lo + vs.size] = xs, vs
    w = np.full(n, L_DOMAIN / n)
    return x, v, w


def host_threads():
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which must not cap the CPU arm)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_baseline(n_s=N_TOTAL, target_seconds=12.0):
    """C restatement of the reference algorithm on the host cores, on the workload's own particle count."""
    orc, nlib = native_oracle()
    threads = host_threads()
    x, v, w = sample_particles(n_s)
    t0 = time.perf_counter()
    orc.baseline_vp_steps(x, v, w, DT, 1, 0.0, L_DOMAIN, N_BASIS, ORDER, 0, threads, nlib)   # warm-up + calibration
    t1 = time.perf_counter() - t0
    steps = int(max(2, min(40, target_seconds / max(t1, 1e-3))))
    t0 = time.perf_counter()
    orc.baseline_vp_steps(x, v, w, DT, steps, 0.0, L_DOMAIN, N_BASIS, ORDER, 0, threads, nlib)
    dt_all = time.perf_counter() - t0
    # single thread (the reference itself is single-threaded), smaller sample
    n_1 = 500_000
    t0 = time.perf_counter()
    orc.baseline_vp_steps(x[:n_1].copy(), v[:n_1].copy(), w[:n_1].copy(), DT, 2, 0.0, L_DOMAIN, N_BASIS, ORDER, 0, 1, nlib)
    dt_1 = time.perf_counter() - t0
    return {"value": n_s * steps / dt_all, "unit": "particle-steps/s", "cores": threads, "kind": "port",
            "sample": f"{n_s} particles (the full workload) x {steps} Strang steps (2 deposits+2 solves+2 gathers each, as the "
                      f"reference), OpenMP {threads} threads, C restatement of the reference algorithm (Julia unavailable)",
            "single_thread_value": n_1 * 2 / dt_1}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    orc, nlib = native_oracle()
    threads = host_threads()
    n_s = args.particles                         # the workload itself: all particles, every step
    x, v, w = sample_particles(n_s)
    for _ in range(args.warmup):
        orc.baseline_vp_steps(x, v, w, DT, 1, 0.0, L_DOMAIN, N_BASIS, ORDER, 0, threads, nlib)
    t0 = time.perf_counter()
    orc.baseline_vp_steps(x, v, w, DT, args.steps, 0.0, L_DOMAIN, N_BASIS, ORDER, 0, threads, nlib)
    dt = time.perf_counter() - t0
    val = n_s * args.steps / dt
    sample = (f"all {n_s} particles of the workload per step, OpenMP {threads} threads; "
              "C restatement of the reference algorithm (Julia toolchain unavailable)")
    cfg = workload_config(args.gpus)
    cfg["particles_total"] = n_s
    emit(({
        "impl": "reference", "metric": "particle-steps/sec", "value": val, "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": val, "unit": "particle-steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(ngpus):
    return {"workload": "bump_on_tail_1d1v_1e8", "particles_total": N_TOTAL, "n_basis": N_BASIS, "spline_order": ORDER,
            "dt": DT, "domain_length": L_DOMAIN, "field_source": "state", "deposit": "deterministic",
            "sharding": f"particles/{ngpus}", "l2": "inputs_larger_than_L2"}


# ------------------------------------------------------------------ GPU arm --
def device_ceilings():
    """tools/microbench/peaks (built by __graft_entry__.build()): fp64 FMA rate, conflict-free shared-memory
    read-modify-write wavefront rate, copy bandwidth of THIS GPU -- the non-HBM ceilings of the particle passes."""
    import subprocess
    exe = ROOT / "tools" / "microbench" / "peaks"
    if not exe.exists():
        return {"error": "tools/microbench/peaks not built"}
    try:
        r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as exc:
        return {"error": repr(exc)}


def parity_check(vm, ctx, rank, world, dist):
    """A small sharded run on these very ranks against the CPU oracle (checker only): 40 001 particles, 6 fused steps
    + diagnostics on meshes of 16 and 256 cells, a CLB right-hand side; phi compared bitwise across ranks."""
    from oracle import vm_oracle as orc
    import torch
    rng = np.random.default_rng(5)
    k, npart, dt, nt = 4, 40001, 0.1, 6
    a, b = 0.0, L_DOMAIN
    x = rng.uniform(a, b, npart); v = rng.standard_normal(npart); w = np.full(npart, (b - a) / npart)
    lo, hi = vm.shard_bounds(npart, rank, world)
    out = {"particles": npart, "steps": nt, "max_rel": 0.0, "ranks_bitwise": True, "meshes": [16, 256]}
    p = vm.DeviceParticles(ctx, hi - lo)
    for n in out["meshes"]:
        fld = vm.DeviceField(ctx, a, b, k, n, 0)
        p.upload(x[lo:hi], v[lo:hi], w[lo:hi])
        diag = fld.run(p, dt, nt, 2, 0, 1.0)
        xg, vg, _ = p.download(w=False)
        phi = fld.coefficients
        err = 0.0
        if rank == 0 or world <= 8:      # every rank checks its own shard against the (replicated) oracle run
            S = orc.periodic_stiffness(a, b, n, k, 0)
            xo, vo = x.copy(), v.copy()
            dref, phiref = orc.integrate_vp(xo, vo, w, dt, 1.0, nt, 2, a, b, n, k, 0, S, want_phi=True)
            err = max(float(np.max(np.abs(xg - xo[lo:hi])) / (b - a)), float(np.max(np.abs(vg - vo[lo:hi])) / np.max(np.abs(vo))),
                      float(np.max(np.abs(diag[:, :3] - dref) / np.max(np.abs(dref), axis=0))),
                      float(np.max(np.abs(phi - phiref[-1])) / np.max(np.abs(phiref[-1]))))
        same = True
        if world > 1:
            t = torch.from_numpy(phi.copy()); ref = t.clone(); dist.broadcast(ref, src=0)
            same = bool(torch.equal(t, ref))
            red = torch.tensor([err, 0.0 if same else 1.0], dtype=torch.float64)
            dist.all_reduce(red, op=dist.ReduceOp.MAX)
            err, same = float(red[0]), red[1] == 0.0
        out["max_rel"] = max(out["max_rel"], err)
        out["ranks_bitwise"] = bool(out["ranks_bitwise"] and same)
        fld.close()
    # order-independent fixed-point deposit (VM_DEPOSIT_FIXED): the sharded run must have the BITS of a one-GPU run of
    # the whole problem (rank 0 repeats it alone on a context without communicator)
    if world == 1 or ctx.peer_connected():
        fld = vm.DeviceField(ctx, a, b, k, 16, 0)
        p.upload(x[lo:hi], v[lo:hi], w[lo:hi])
        fld.run(p, dt, 4, 0, vm._lib.VM_RUN_FIXED_DEPOSIT, 1.0)
        phis = fld.coefficients.copy()
        solo = vm.Context(ctx.device)
        f1 = vm.DeviceField(solo, a, b, k, 16, 0)
        p1 = vm.DeviceParticles(solo, npart)
        p1.upload(x, v, w)
        solo.set_tuning("bankq", 1)                # ... and with the other deposit layout
        f1.run(p1, dt, 4, 0, vm._lib.VM_RUN_FIXED_DEPOSIT, 1.0)
        same = bool(np.array_equal(f1.coefficients, phis))
        if world > 1:
            red = torch.tensor([0.0 if same else 1.0], dtype=torch.float64)
            dist.all_reduce(red, op=dist.ReduceOp.MAX)
            same = red[0] == 0.0
        out["fixed_point_bits_equal_single_gpu_run"] = bool(same)
        solo.close(); fld.close()
    # v-space: sharded conservative Lenard-Bernstein right-hand side
    vs = vm.DeviceVSpline(ctx, -10.0, 10.0, 41, 4, 1)
    wv = np.full(npart, 1.0 / npart)
    p.upload(v=v[lo:hi], w=wv[lo:hi])
    vdot = vs.lb_rhs(p, 1.0, True)
    M = orc.dirichlet_mass(-10.0, 10.0, 41, 4)
    vref, _, _ = orc.lb_rhs(v, wv, -10.0, 10.0, 41, 4, M, 1.0, True)
    e = float(np.max(np.abs(vdot - vref[lo:hi])) / np.max(np.abs(vref)))
    if world > 1:
        red = torch.tensor([e], dtype=torch.float64)
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        e = float(red[0])
    out["clb_rhs_max_rel"] = e
    out["tolerance"] = 1e-10
    out["ok"] = bool(out["max_rel"] <= 1e-10 and e <= 1e-10 and out["ranks_bitwise"]
                     and out.get("fixed_point_bits_equal_single_gpu_run", True))
    vs.close(); p.close()
    return out


def run_gpu(args):
    from __graft_entry__ import load_package
    vm = load_package()
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    try:        # one core set per rank: pinned host buffers are first touched (and stay) near the cores that fill them
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // max(world, 1))
        os.sched_setaffinity(0, set(cores[local * per:(local + 1) * per]) or set(cores))
    except Exception:
        pass
    if world > 1:
        dist.init_process_group(backend="gloo")      # host-side rendezvous/barrier only
    ctx = vm.init_distributed_context(local, peer_exchange=not args.no_peer)
    if args.no_fuse:
        ctx.set_tuning("no_fuse", 1)
    if args.no_pdl:
        ctx.set_tuning("no_pdl", 1)
    ntot = args.particles
    lo, hi = vm.shard_bounds(ntot, rank, world)
    nloc = hi - lo

    def barrier():
        ctx.sync()
        if world > 1:
            dist.barrier()

    def over_ranks(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=op)
        return float(t[0])

    def max_over_ranks(x):
        return over_ranks(x, dist.ReduceOp.MAX) if world > 1 else x

    def min_over_ranks(x):
        return over_ranks(x, dist.ReduceOp.MIN) if world > 1 else x

    parity = None
    if not args.no_parity:
        try:
            parity = parity_check(vm, ctx, rank, world, dist)
        except Exception as exc:
            parity = {"ok": False, "error": repr(exc)}

    fld = vm.DeviceField(ctx, 0.0, L_DOMAIN, ORDER, N_BASIS, 0)
    p = vm.DeviceParticles(ctx, nloc)
    p.fill(vm._lib.VM_FILL_BUMP_ON_TAIL, [EPS, KAPPA, ALPHA, SIGMA, V0], SEED, lo, ntot)
    flags = vm._lib.VM_RUN_ATOMIC_DEPOSIT if args.atomic else 0
    if args.general_weights:
        ctx.set_tuning("no_uniform_w", 1)
    ALG_BYTES_PER_PARTICLE = ALG_BYTES_GENERAL_W if args.general_weights else ALG_BYTES_UNIFORM_W
    dep_bytes = ALG_BYTES_PER_PARTICLE - 24      # deposit-only pass: read x (+ w)
    peak, peak_src = peaks()

    # ---------------- device-resident timing ----------------
    # Region A (value): K steps, nothing but the hot path on the stream; `repeats` such regions, median reported.
    # Before each region one untimed fused step aligns the ranks ON THE DEVICE (its exchange waits for the slowest
    # rank): a host barrier alone leaves the ranks' streams hundreds of microseconds apart, which the first in-kernel
    # exchange of a 2 ms region would charge to the earliest rank.
    def timed_region(field, steps, e0, e1):
        barrier()
        field.run(p, DT, 1, 0, flags, 1.0)           # untimed: device-side alignment of the ranks
        lc = ctx.launch_count()
        ctx.event_record(e0)
        field.run(p, DT, steps, 0, flags, 1.0)
        ctx.event_record(e1)
        lc = ctx.launch_count() - lc
        barrier()
        loc = ctx.event_elapsed_ms(e0, e1)
        return max_over_ranks(loc), min_over_ranks(loc), lc

    fld.run(p, DT, args.warmup, 0, flags, 1.0)
    sampler = ClockSampler(physical_gpu_index(local))
    barrier()
    sampler.start()
    regions = [timed_region(fld, args.steps, 0, 1) for _ in range(args.repeats)]
    clocks = sampler.stop()
    launches = regions[0][2]                                     # kernels launched inside ONE timed region of K steps
    ms_all = sorted(r[0] for r in regions)
    ms = float(np.median(ms_all))
    value = ntot * args.steps / (ms * 1e-3)
    timing = {"repeats": args.repeats, "ms_per_step_median": ms / args.steps, "ms_per_step_min": ms_all[0] / args.steps,
              "ms_per_step_max": ms_all[-1] / args.steps,
              "rank_min_over_max_elapsed": min(r[1] / r[0] for r in regions),
              "aligned_by": "one untimed fused step after the host barrier"}

    # Region B (roofline): the same K steps again with every launch of the dominant kernel bracketed by
    # CUDA events on the library's stream (the brackets serialise the launches, so they are kept out of
    # region A where programmatic dependent launch overlaps consecutive kernels).
    def bracketed(field, steps, alg_bytes, kernel):
        ctx.set_tuning("profile", 1)
        ctx.profile_read()
        barrier()
        ctx.event_record(6)
        field.run(p, DT, steps, 0, flags, 1.0)
        ctx.event_record(7)
        barrier()
        ms_b = max_over_ranks(ctx.event_elapsed_ms(6, 7))
        kn, kms = ctx.profile_read()
        ctx.set_tuning("profile", 0)
        if kn <= 0:
            return None
        achieved = alg_bytes * nloc / (kms / kn * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": f"of {peak_src}", "launches_timed": kn, "avg_launch_ms": kms / kn,
                "ms_per_step_with_brackets": ms_b / steps, "algorithmic_bytes_per_launch": alg_bytes * nloc,
                "algorithmic_bytes_per_particle": alg_bytes}

    def traffic_of(general):
        tf = ROOT / "profiles" / "traffic.json"
        try:
            key = "k_vp_pass_push_deposit_bytes_per_particle" + ("" if general else "_uniform_w")
            t = json.loads(tf.read_text()).get(key)
            return t * nloc if t is not None else None
        except Exception:
            return None

    roofline = bracketed(fld, args.steps, ALG_BYTES_PER_PARTICLE, "k_vp_pass<4,PRIV,PUSH_DEPOSIT>")
    if roofline:
        roofline["traffic"] = traffic_of(args.general_weights)
        roofline["traffic_source"] = "ncu --set full capture of this kernel (profiles/traffic.json), not re-measured in this run"
        roofline["weights"] = ("per-particle array streamed" if args.general_weights else
                               "uniform (w = L/N for every particle, as every sampler of the reference produces): passed as a kernel parameter, not read from HBM")

    # the same step with the weight array streamed (SURVEY 8(d)'s 40 B/particle-step), beside the 32 B headline
    general = None
    if not args.general_weights:
        ctx.set_tuning("no_uniform_w", 1)
        fld.run(p, DT, 3, 0, flags, 1.0)
        g_all = sorted(timed_region(fld, args.steps, 4, 5)[0] for _ in range(3))
        g_ms = float(np.median(g_all))
        g_roof = bracketed(fld, args.steps, ALG_BYTES_GENERAL_W, "k_vp_pass<4,PRIV,PUSH_DEPOSIT> (weights streamed)")
        if g_roof:
            g_roof["traffic"] = traffic_of(True)
        general = {"algorithmic_bytes_per_particle": ALG_BYTES_GENERAL_W, "ms_per_step": g_ms / args.steps,
                   "particle_steps_per_s": ntot * args.steps / (g_ms * 1e-3),
                   "step_hbm_frac": ALG_BYTES_GENERAL_W * nloc * args.steps / (g_ms * 1e-3) / 1e9 / peak, "roofline": g_roof}
        ctx.set_tuning("no_uniform_w", 0)

    # ---------------- end-to-end through host buffers ----------------
    hx = torch.empty(nloc, dtype=torch.float64).pin_memory()
    hv = torch.empty(nloc, dtype=torch.float64).pin_memory()
    p.fill(vm._lib.VM_FILL_BUMP_ON_TAIL, [EPS, KAPPA, ALPHA, SIGMA, V0], SEED, lo, ntot)
    p.download(w=False, out=(hx.numpy(), hv.numpy(), None))
    e2e_steps = args.steps
    w0 = L_DOMAIN / ntot
    hw = None
    if args.general_weights:
        hw = torch.full((nloc,), w0, dtype=torch.float64).pin_memory()

    def e2e_once():
        if hw is None:
            p.upload(hx.numpy(), hv.numpy(), None)                       # H2D 16 B/particle
            p.set_uniform_weight(w0)                                     # w = L/N declared, not uploaded
        else:
            p.upload(hx.numpy(), hv.numpy(), hw.numpy())                 # H2D 24 B/particle
        d = fld.run(p, DT, e2e_steps, e2e_steps, flags, 1.0)             # K steps + [W,K,M] rows read back
        p.download(w=False, out=(hx.numpy(), hv.numpy(), None))          # D2H 16 B/particle
        return d

    e2e_once()                                                           # warm-up (page-locks, allocations)
    p.fill(vm._lib.VM_FILL_BUMP_ON_TAIL, [EPS, KAPPA, ALPHA, SIGMA, V0], SEED, lo, ntot)
    p.download(w=False, out=(hx.numpy(), hv.numpy(), None))
    barrier()
    t0 = time.perf_counter()
    ctx.event_record(2)
    diag = e2e_once()
    ctx.event_record(3)
    barrier()
    wall = time.perf_counter() - t0
    e2e_ms = max_over_ranks(max(ctx.event_elapsed_ms(2, 3), 0.0))
    h2d = (24 if hw is not None else 16) * ntot
    e2e = {"value": ntot * e2e_steps / (e2e_ms * 1e-3), "unit": "particle-steps/s",
           "h2d_bytes_per_step": h2d / e2e_steps, "d2h_bytes_per_step": (16 * ntot + 2 * 32 * world) / e2e_steps,
           "steps_per_call": e2e_steps, "ms_per_call": e2e_ms, "wall_ms": wall * 1e3,
           "what": "upload x,v from pinned host arrays (+ vm_particles_set_uniform_weight: w = L/N is declared, not "
                   "uploaded) + K fused steps + diagnostics read-back + download x,v "
                   "(vm_particles_upload_soa / vm_vp_run / vm_particles_download_soa)",
           "energy_drift": float(abs((diag[-1, 0] + diag[-1, 1]) - (diag[0, 0] + diag[0, 1])) / (diag[0, 0] + diag[0, 1]))}

    # ---------------- secondary numbers (BASELINE metric: "+ deposit HBM GB/s vs peak"; configs[2], [3]) ----------
    def timed(fn, reps):
        fn(); ctx.sync(); ctx.event_record(8)
        for _ in range(reps):
            fn()
        ctx.event_record(9)
        return max_over_ranks(ctx.event_elapsed_ms(8, 9)) / reps

    ceil = device_ceilings() if (rank == 0 and not args.no_ceilings) else None
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    sm_count = ctx.device_info()["sm_count"]

    def lsu_frac(wavefronts_per_warp_particle, ms_):
        """Shared-memory data-pipe utilisation implied by the wavefront model (profiles/README.md section 7) at the
        SM clock sampled in region A, against the measured conflict-free read-modify-write rate of this GPU."""
        if not ceil or "smem_rmw_wavefronts_per_clk_per_sm" not in ceil:
            return None
        wf = wavefronts_per_warp_particle * (nloc / 32.0) / sm_count
        rate = wf / (ms_ * 1e-3 * sm_mhz * 1e6)
        return {"wavefronts_per_warp_of_particles": wavefronts_per_warp_particle, "achieved_wf_per_clk_per_sm": rate,
                "peak_wf_per_clk_per_sm": ceil["smem_rmw_wavefronts_per_clk_per_sm"],
                "frac": rate / ceil["smem_rmw_wavefronts_per_clk_per_sm"], "sm_mhz_assumed": sm_mhz}

    dep_ms = timed(lambda: fld.deposit(p, 0), 10)            # projection!(potential, dist) alone
    ctx.set_tuning("no_uniform_w", 1)                        # same pass with the weight array streamed (16 B/particle)
    dep_ms_gw = timed(lambda: fld.deposit(p, 0), 10)
    ctx.set_tuning("no_uniform_w", 1 if args.general_weights else 0)
    deposit = {"kernel": "k_vp_pass<4,PRIV,DEPOSIT>", "ms": dep_ms, "GBps": dep_bytes * nloc / dep_ms / 1e6,
               "frac": dep_bytes * nloc / dep_ms / 1e6 / peak, "algorithmic_bytes_per_particle": dep_bytes,
               "bound": "lsu" if dep_bytes == 8 else "hbm", "lsu": lsu_frac(18.0, dep_ms),
               "particles_per_s": nloc * world / dep_ms * 1e3,
               "per_particle_weights": {"ms": dep_ms_gw, "algorithmic_bytes_per_particle": 16, "bound": "hbm",
                                        "GBps": 16 * nloc / dep_ms_gw / 1e6, "frac": 16 * nloc / dep_ms_gw / 1e6 / peak},
               "note": "with the uniform weight of this workload the pass reads 8 B/particle and is bound by the L1TEX/"
                       "shared-memory data pipe, not by HBM: 16 wavefronts per warp of particles for the K=4 read-modify-"
                       "writes of the lane-private replicas + 2 for the loads (`lsu`: against the measured wavefront rate "
                       "of tools/microbench/peaks); with per-particle weights (16 B/particle) see per_particle_weights",
               "shared_atomics": 0, "mode": "deterministic (lane-private replicas)"}
    secondary = None
    if not args.no_secondary:
        vs = vm.DeviceVSpline(ctx, -10.0, 10.0, 41, 4, 1)
        p.fill(vm._lib.VM_FILL_DOUBLE_MAXWELLIAN, [-10.0, 10.0, 2.0], SEED, lo, ntot)
        lb = timed(lambda: vs.lb_rhs(p, 1.0, False, to_host=False), 5)
        clb = timed(lambda: vs.lb_rhs(p, 1.0, True, to_host=False), 5)
        rk = timed(lambda: vs.rk438_run(p, 1e-3, 5, 1.0, True, 0), 2) / 5
        # mesh-size sweep of the fused step (BASELINE configs[4]: 64-1024 spline modes), same particles, cubic
        p.fill(vm._lib.VM_FILL_BUMP_ON_TAIL, [EPS, KAPPA, ALPHA, SIGMA, V0], SEED, lo, ntot)
        mesh = {}
        for nh in ((32, 64, 128, 256, 512, 1024) if world == 1 else (64, 256, 1024)):
            f2 = vm.DeviceField(ctx, 0.0, L_DOMAIN, ORDER, nh, 0)
            f2.run(p, DT, 3, 0, flags, 1.0)
            t = float(np.median([timed_region(f2, 10, 10, 11)[0] for _ in range(3)])) / 10
            plan = vm._lib.pass_plan(nh, ORDER, 1)
            layout = {0: "lane-private replicas", 4: "bank-sorted queues",
                      5: "limb atomics, %d bank-steered replicas, %d-fold gather table" % (plan.replicas, plan.gather_copies)}.get(plan.variant, str(plan.variant))
            mesh[str(nh)] = {"ms_per_step": t, "particle_steps_per_s": ntot / t * 1e3,
                             "step_hbm_frac": ALG_BYTES_PER_PARTICLE * nloc / t / 1e6 / peak, "deposit_layout": layout}
            f2.close()
        secondary = {"particles_total": ntot, "vspline": "41 knots, order 4, Dirichlet, v in (-10,10)",
                     "mesh_sweep_n_basis": mesh,
                     "weights": "uniform w = 1/N: not streamed (8 B less per deposit pass)",
                     "lb_rhs_evals_per_s": ntot / lb * 1e3, "lb_rhs_hbm_frac": 24 * nloc / lb / 1e6 / peak,
                     "clb_rhs_evals_per_s": ntot / clb * 1e3, "clb_rhs_hbm_frac": 32 * nloc / clb / 1e6 / peak,
                     "clb_rk438_particle_steps_per_s": ntot / rk * 1e3, "clb_rk438_hbm_frac": 168 * nloc / rk / 1e6 / peak,
                     "bound_note": "the LB/CLB right-hand sides move 24/32 B per particle through three passes that are "
                                   "co-limited by the shared-memory data pipe (lane-private v-deposit: 18 wavefronts per warp "
                                   "of particles) and instruction issue, see profiles/README.md"}
        vs.close()

    cpu = None
    if world == 1 and rank == 0 and not args.no_cpu:
        p.close()                                  # the CPU leg needs the host's memory bandwidth, not the GPU's
        try:
            cpu = cpu_baseline(ntot)
        except Exception as exc:   # the CPU leg must never take the GPU number down with it
            cpu = {"error": repr(exc)}

    if rank == 0:
        cfg = workload_config(world)
        cfg["particles_total"] = ntot
        cfg["collective"] = ("none" if world == 1 else
                             "nccl_allreduce" if (args.no_peer or args.no_fuse or world > 8) else
                             "fused_nvlink_peer_exchange_in_deposit_kernel")
        emit(({
            "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "roofline": roofline, "general_weights": general, "deposit": deposit, "secondary": secondary,
            "cpu_baseline": cpu, "e2e": e2e, "parity": parity, "timing": timing, "ceilings": ceil,
            "gpu_launches": int(launches), "clocks": clocks,
            "step_hbm_frac": ALG_BYTES_PER_PARTICLE * nloc * args.steps / (ms * 1e-3) / 1e9 / peak,
            # the same step against the copy kernel of tools/microbench/peaks timed IN THIS RUN (same box, same power /
            # clock state: under the pool's sw_power_cap the copy kernel itself drops from ~6.5 to ~6.0 TB/s)
            "step_frac_of_in_run_copy_rate": (ALG_BYTES_PER_PARTICLE * nloc * args.steps / (ms * 1e-3) / 1e9 / ceil["hbm_copy_gbs"]
                                               if isinstance(ceil, dict) and ceil.get("hbm_copy_gbs") else None),
        }))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def emit(obj):
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, torchrun notes) was
    redirected to stderr at start-up."""
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


_REAL_STDOUT = os.dup(1)


def main():
    os.dup2(2, 1)          # libraries that print to fd 1 (NCCL version banner) must not corrupt the JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", type=int, default=N_TOTAL)
    ap.add_argument("--atomic", action="store_true", help="use the shared-atomic deposit variant (A/B)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--general-weights", action="store_true", help="stream the per-particle weight array even though it is uniform (A/B: 40 B/particle)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the Lenard-Bernstein secondary numbers")
    ap.add_argument("--no-peer", action="store_true", help="NCCL all-reduce instead of the fused NVLink peer-memory exchange (A/B)")
    ap.add_argument("--no-pdl", action="store_true", help="disable programmatic dependent launch of the pass kernels (A/B)")
    ap.add_argument("--no-fuse", action="store_true", help="separate reduce/solve kernels instead of the last-CTA finish (A/B)")
    ap.add_argument("--repeats", type=int, default=5, help="timed regions of K steps each; the median is reported")
    ap.add_argument("--no-parity", action="store_true", help="skip the small sharded run checked against the CPU oracle")
    ap.add_argument("--no-ceilings", action="store_true", help="skip tools/microbench/peaks")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    args.repeats = max(args.repeats, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
