#!/usr/bin/env python
"""bench.py -- particle-updates/s of the MFV step (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A "step" is one full pass of the hot path (MeshlessScheme::run loop body, phases 1-16: cell sort ->
neighbour search -> density/matrices -> gradients/limiter/CFL -> faces + exact Riemann -> update) over
all particles of the workload.  Default workload at EVERY N = BASELINE.json configs[3]: Kelvin-Helmholtz
2D MFV, N = 2000^2 = 4 M particles (the largest configuration that BASELINE lists for one GPU, "at
1/2/4/8 B200"); with N>1 the same 4 M particles are slab-decomposed over the GPUs (NCCL halo exchange +
allreduce-min of dt) -> STRONG scaling.  `--workload sedov61` = configs[1] (Sedov 3D 61^3; with N>1 weak
scaling, ~61^3 per GPU), `--workload sedov256` = configs[4].  At N=8 the default run appends the
Sedov 256^3 numbers (configs[4]) under "also" unless --no-also is given.  Initial conditions are synthetic
(jittered lattices of those shapes, seed 6102003; SEAGen/HDF5 files are not available).

The timed region is made at least 0.5 s long (so that the 100 ms clock sampler covers it): the block of
--steps steps is repeated `timed_region.blocks` times back to back inside ONE region, ms_per_step is the
mean over all of them.  With N>1 rank 0 also runs a small case sharded and on one GPU and reports
whether the results agree bit for bit ("mgpu_check").

One JSON line on stdout (rank 0).  `value` = device-resident throughput (CUDA events on the library's
stream), `e2e` = the same metric through the C ABI with HOST (pinned) buffers: upload + step + download
every step.  `roofline` describes the dominant kernel of the per-kernel profile,
`cpu_baseline` the reference's own sources (oracle/_ref) timed on one host core on a bounded sample.
`--impl reference` times only that CPU reference arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle_updates_per_s"
UNIT = "particle-updates/s"

# ---- algorithmic work figures ----
ALL_ALGO_BYTES = {2: 508, 3: 788}  # bytes per particle-update, whole step


def workloads():
    from meshlesshydro_b200 import ic as IC
    return {
        "sedov61": (lambda n=61: IC.sedov(n), "sedov3d", "Sedov blast 3D MFV, N=61^3 jittered lattice (BASELINE configs[1])"),
        "sedov128": (lambda: IC.sedov(128), "sedov3d", "Sedov blast 3D MFV, N=128^3"),
        "sedov256": (lambda: IC.sedov(256), "sedov3d", "Sedov blast 3D MFV, N=256^3 (BASELINE configs[4])"),
        "kh100": (lambda: IC.kelvin_helmholtz(100, lattice=False), "kh2d", "Kelvin-Helmholtz 2D MFV, N=10^4 random (BASELINE configs[0] shape)"),
        "kh1000": (lambda: IC.kelvin_helmholtz(1000, lattice=True, jitter=0.2), "kh2d", "Kelvin-Helmholtz 2D MFV, N=1M jittered lattice"),
        "kh2000": (lambda: IC.kelvin_helmholtz(2000, lattice=True, jitter=0.2), "kh2d", "Kelvin-Helmholtz 2D MFV, N=4M jittered lattice (BASELINE configs[3])"),
        "fb1000": (lambda: IC.fluid_block(1000, jitter=0.05), "fb2d", "fluid-block 2D, N=10^6 (BASELINE configs[2])"),
    }


# ------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks line")
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [t.strip() for t in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [v for v in sm if v > 0.5 * (max(sm) if sm else 0)]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU reference arm: the reference's own sources (oracle/_ref), one host core (the reference is serial)
# ------------------------------------------------------------------------------------------------
def _ref_variant(preset):
    return {"sedov3d": "sedov3d_bench", "kh2d": "kh2d_bench", "fb2d": "fb2d"}[preset]


def time_reference(ic_factory, preset, n_side, steps, warmup, budget_s):
    """Times `steps` steps of the reference-source build on a bounded sample: the workload's IC generator
    at side length n_side, shrunk if one step would blow the time budget.  Returns dict."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from cpu_oracles import Reference, Oracle, make_config
    from meshlesshydro_b200 import ic as IC
    variant = _ref_variant(preset)
    gen = {"sedov3d": IC.sedov, "kh2d": lambda n: IC.kelvin_helmholtz(n, lattice=True, jitter=0.2),
           "fb2d": lambda n: IC.fluid_block(n, jitter=0.05)}[preset]
    dim = 3 if preset == "sedov3d" else 2

    def per_particle_s(side):
        # measured in this image (one core): 3D 3.2e-5 s; 2D 4.0e-5 s, plus the reference's brute-force ghost search
        # (Particles::ghostNNS, O(N N_ghost)) in periodic runs: 4.4e-5 at side 150, 5.1e-5 at side 300
        return 3.2e-5 if dim == 3 else (4.0e-5 + 3.6e-8 * side if preset == "kh2d" else 4.0e-5)
    kind = "reference" if Reference.available(variant) else "port"
    while True:
        n = n_side ** dim
        est = per_particle_s(n_side) * n * (steps + warmup)
        if est <= budget_s or n_side <= 16:
            break
        n_side = max(16, min(n_side - 1, int(n_side * (budget_s / est) ** (1.0 / dim))))
    ic = gen(n_side)
    n = len(ic["x"])
    if kind == "reference":
        arm = Reference(variant, ic)
    else:
        arm = Oracle(make_config(preset, ic["h"], ic["gamma"], ic.get("box"), abs_mode=0), ic)
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    for _ in range(warmup):
        arm.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        arm.step()
    dt = time.perf_counter() - t0
    arm.close()
    return {"value": n * steps / dt, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "%d steps (+%d warm-up) of the same IC generator at N=%d (%s, side %d), %s, %.1f s"
                      % (steps, warmup, n, preset, n_side,
                         "oracle/_ref = reference sources g++ -std=c++11 -O3" if kind == "reference" else "oracle/liboracle.so port",
                         dt),
            "ms_per_step": 1e3 * dt / steps, "n_particles": n}


# ------------------------------------------------------------------------------------------------
SIDES = {"sedov61": 61, "sedov128": 128, "sedov256": 256, "kh100": 100, "kh1000": 1000, "kh2000": 2000, "fb1000": 1000}
BLOCK_CAP = {"fb1000": 5}  # steps per timed block (see run_workload)
# cpu_baseline sample: 10-30 s of single-core work (time_reference shrinks the side to fit its 25 s budget)
CPU_SIDES = {"sedov61": 61, "sedov128": 61, "sedov256": 61, "kh100": 100, "kh1000": 400, "kh2000": 400, "fb1000": 400}


def mgpu_bitwise_check(dist, local_rank, rank, world):
    """A small periodic case run sharded over all ranks and, on rank 0, on one GPU: states and time steps must agree
    bit for bit (same neighbour order, same arithmetic; cut faces are evaluated identically on both ranks)."""
    from meshlesshydro_b200 import capi, multigpu, ic as IC
    ic = IC.kelvin_helmholtz(max(128, 32 * world), lattice=True, jitter=0.2)
    steps = 3

    def cfg_():
        c = capi.make_config("kh2d", ic["h"], ic["gamma"], ic.get("box"), abs_mode=capi.ABS_INT_TRUNC, max_interactions=96)
        c.device = local_rank
        return c
    gpu, _ = multigpu.create_sharded(cfg_(), ic, dist)
    dts = [gpu.step() for _ in range(steps)]
    st = gpu.download_state()
    flags = gpu.error_flags()
    gpu.close()
    names = ["x", "y", "vx", "vy", "m", "u"]
    gathered = [None] * world
    dist.all_gather_object(gathered, {k: st[k] for k in names + ["ids"]})
    res = None
    if rank == 0:
        N = len(ic["x"])
        full = {k: np.full(N, np.nan) for k in names}
        count = np.zeros(N, dtype=np.int64)
        for part in gathered:
            count[part["ids"]] += 1
            for k in names:
                full[k][part["ids"]] = part[k]
        one = capi.MfvGpu(cfg_())
        one.upload(ic)
        dts1 = [one.step() for _ in range(steps)]
        ref = one.download_state()
        one.close()
        equal = bool(np.all(count == 1)) and dts == dts1 and all(np.array_equal(full[k], ref[k]) for k in names)
        res = {"case": "KH 2D N=%d periodic, %d steps, %d slabs vs 1 GPU" % (N, steps, world), "bitwise_equal": equal,
               "ownership_is_partition": bool(np.all(count == 1)), "device_flags": flags}
    return res


def run_workload(args, wname, dist, rank, world, local_rank, want_profile=True, want_e2e=True):
    """Time one workload on `world` GPUs.  Returns the measurement dict (rank 0) or None."""
    import torch
    from meshlesshydro_b200 import capi
    factory, preset, wdesc = workloads()[wname]
    scaling = "strong"
    n_side_weak = None
    if world > 1 and wname == "sedov61":
        scaling = "weak"
        n_side_weak = int(round(61 * world ** (1.0 / 3.0)))
        wdesc = "Sedov blast 3D MFV, weak scaling: N=%d^3 over %d slabs (~61^3 per GPU)" % (n_side_weak, world)
    ic = factory(n_side_weak) if n_side_weak else factory()
    D = ic["dim"]
    n_total = len(ic["x"])
    cfg = capi.make_config(preset, ic["h"], ic["gamma"], ic.get("box"), abs_mode=capi.ABS_INT_TRUNC,
                           q13_mode=capi.Q13_ZERO_Z, max_interactions=128 if D == 3 else 96)
    cfg.device = local_rank
    cfg.rank, cfg.nranks = rank, world
    ids_local = None
    if world > 1:
        from meshlesshydro_b200 import multigpu
        gpu, ic_local = multigpu.create_sharded(cfg, ic, dist)
        ids_local = multigpu.shard(ic, rank, world)[1]
    else:
        gpu, ic_local = capi.MfvGpu(cfg), ic
        gpu.upload(ic_local)
    n_local = len(ic_local["x"])

    def barrier():
        gpu.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ----
    # The timed region is `blocks` blocks of --steps steps.  Every block starts from the initial condition (re-uploaded,
    # plus one step, both untimed): the reference's scheme is not stable on the Sedov blast -- its CPU restatement turns
    # NaN near t = 0.015 (tests/test_gpu_fullrun.py) -- so a benchmark that simply kept stepping for 0.6 s would leave the
    # regime in which the step does meaningful work.  ms = sum of the blocks' CUDA-event times (max over ranks per block).
    def restart():
        gpu.upload(ic_local, ids=ids_local)
        gpu.step(want_dt=False)

    for _ in range(args.warmup):
        gpu.step(want_dt=False)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    # fluid block: the free surface of the reference's scheme breaks up within ~10 steps at this resolution (vacuum
    # faces, clustering until the lists overflow -- device flags 35, the reference exits there); blocks of <= 5 steps
    bsteps = min(args.steps, BLOCK_CAP.get(wname, args.steps))
    ms, launches, blocks = 0.0, 0, 0
    while (ms < 1e3 * args.min_seconds or blocks < 1) and blocks < 400:  # >= 0.6 s of timed steps (ms = max over ranks: same on every rank)
        restart()
        l0 = gpu.launch_count()
        barrier()
        gpu.timer_start()
        for _ in range(bsteps):
            gpu.step(want_dt=False)
        ms_b = gpu.timer_stop()
        barrier()
        launches += gpu.launch_count() - l0  # kernels launched inside the timed blocks
        ms += max_over_ranks(ms_b)
        blocks += 1
    clocks = sampler.stop() if rank == 0 else None
    flags = gpu.error_flags()
    nsteps = bsteps * blocks
    value = n_total * nsteps / (ms * 1e-3)
    res = {"wname": wname, "wdesc": wdesc, "scaling": scaling, "D": D, "n_total": n_total, "n_local": n_local, "value": value,
           "ms_per_step": ms / nsteps,
           "timed_region": {"blocks": blocks, "steps_total": nsteps, "seconds": ms * 1e-3,
                            "steps_per_block": bsteps, "note": "every block of --steps steps starts from the re-uploaded initial condition (+1 step), untimed"},
           "launches": launches, "clocks": clocks, "flags": flags, "h": ic["h"]}

    # ---- per-kernel profile (separate, untimed pass) for the roofline object ----
    if want_profile:
        restart()
        gpu.profile(True)
        psteps = max(3, min(bsteps, 10))
        for _ in range(psteps):
            gpu.step(want_dt=False)
        prof = gpu.profile_read()
        gpu.profile(False)
        res["prof"] = prof
        res["psteps"] = psteps
        res["noi_mean"] = float((gpu.fetch("noi").mean() + gpu.fetch("noiGhosts").mean()))  # rank 0's particles when sharded
        res["nfaces"] = int(gpu.fetch("num_faces")[0])

    # ---- end to end through the C ABI with host buffers ----
    if want_e2e:
        # every rank: H2D of its shard from pinned host arrays, one step, D2H of the result.  Single GPU: the next step
        # starts from the downloaded state (host round trip).  Sharded: particles migrate between slabs, so the owned set
        # a rank downloads may differ from the one it uploaded; every step therefore uploads the rank's initial shard
        # (same bytes, same work) and the download buffers are sized for the largest owned set.
        names = ["x", "y", "vx", "vy", "m", "u"] + (["z", "vz"] if D == 3 else [])
        host = {k: capi.pinned_empty(n_local) for k in names}
        for k in names:
            host[k][:] = ic_local[k]
        host_ic = dict(ic_local)
        host_ic.update(host)
        n_out = n_local if world == 1 else int(cfg.capacity)
        out = {k: (capi.pinned_empty(n_out) if k in names else None) for k in ["x", "y", "z", "vx", "vy", "vz", "m", "u"]}
        out["ids"] = capi.pinned_empty(n_out, np.int32)
        esteps = max(3, min(bsteps, 10))
        out_ic = dict(ic_local)
        out_ic.update({k: out[k] for k in names})
        for _ in range(2):
            gpu.upload(host_ic, ids=ids_local)
            gpu.step(want_dt=False)
            gpu.download_state(out)
        barrier()
        t0 = time.perf_counter()
        src, dst = host_ic, out
        for _ in range(esteps):
            gpu.upload(src, ids=ids_local)      # H2D of this step's inputs
            gpu.step(want_dt=False)
            gpu.download_state(dst)             # D2H of the step's result (synchronises)
            if world == 1:
                # the next step starts from the state just downloaded (host round trip): the two sets of pinned
                # buffers swap roles, as a host code that owns its particle arrays would do
                if src is host_ic:
                    src, dst = out_ic, {**{k: (host[k] if k in names else None) for k in out if k != "ids"}, "ids": out["ids"]}
                else:
                    src, dst = host_ic, out
        barrier()
        e_s = max_over_ranks(time.perf_counter() - t0)
        nb = len(names) * 8 * n_local
        res["e2e"] = {"value": n_total * esteps / e_s, "unit": UNIT, "h2d_bytes_per_step": nb, "d2h_bytes_per_step": nb + 4 * n_local,
                      "steps": esteps, "ms_per_step": 1e3 * e_s / esteps,
                      "timing": "host wall clock (max over ranks) around upload+step+download, pinned buffers; bytes are rank 0's"}
    gpu.close()
    return res


def rooflines(res, local_rank):
    """roofline object of the dominant kernel + every kernel against both rooflines (DESIGN.md sections 3 and 6)"""
    from meshlesshydro_b200 import capi
    D, n_local, nfaces, noi_mean = res["D"], res["n_local"], res["nfaces"], res["noi_mean"]
    prof, psteps, wname = res["prof"], res["psteps"], res["wname"]
    per_launch = {k: v[0] / psteps for k, v in prof.items() if v[1]}  # per STEP (a chunked kernel may launch more than once)
    step_ms_prof = sum(v[0] for v in prof.values()) / psteps
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "of measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "of fallback (B200_PROFILING.md 6.65 TB/s)"
    fp64_peak = capi.fp64_peak_tflops(local_rank)
    ops, ops_note = {}, None
    try:
        allops = json.load(open(os.path.join(ROOT, "profiles", "fp64_ops.json")))
        ops = allops.get(wname, {})
        if not ops and wname == "kh2000" and "kh1000" in allops:
            # same flow, same neighbour count: per-launch counts scale with the particle count
            ops = {k: {kk: vv * 4.0 for kk, vv in v.items()} for k, v in allops["kh1000"].items()}
            ops_note = "ncu counts of kh1000 (1 M) x 4"
    except Exception:
        pass
    top = max(per_launch.items(), key=lambda kv: prof[kv[0]][0])[0]
    top_ms = per_launch[top]
    nslots = n_local * noi_mean if noi_mean else None
    frec = 4 if D == 2 else 6
    # SURVEY 8(d) algorithmic bytes per launch (each per-particle array once per kernel that must touch it; lists and
    # per-face staging excluded)
    algo = {
        "k1_gather": lambda: n_local * 2 * ((2 * D + 2) * 8 + 8),
        "k4c_flux_sum_update": lambda: n_local * 2 * (2 * D + 2) * 8,
    }
    # bytes the kernel is DESIGNED to move (algorithmic + neighbour lists + per-face staging of the flux pass): what its
    # measured DRAM traffic should be compared with -- not SURVEY's algorithmic figure
    staged = {
        "k4a_face_states": lambda: n_local * (2 * D + 4 + D * D + (D + 2) * D) * 8 + nfaces * ((4 * D + 4) * 8 + 8),
        "k4c_flux_sum_update": lambda: nslots * (4 + frec * 8) + n_local * 2 * (2 * D + 2) * 8,
        "k2b_face_index": lambda: nslots * 12 + nfaces * 8,
        "k4b1_face_setup": lambda: nfaces * (6 * 8 + 12 * 8 + 4),
        "k4b3_face_finish": lambda: nfaces * ((4 * D + 4) * 8 + 8 + frec * 8),
    }
    roofline = {"kernel": top, "share_of_step": prof[top][0] / psteps / step_ms_prof if step_ms_prof else None,
                "ms_per_step": top_ms, "launches_per_step": prof[top][1] / psteps, "traffic": None,
                "counts": "per step of this rank's shard" + ("; " + ops_note if ops_note else "")}
    kops = ops.get(top)
    scale_n = 1.0
    if kops and res.get("world", 1) > 1:
        scale_n = n_local / float(res["n_total"])  # ncu counted the single-GPU workload: this rank holds n_local of it
    if kops:
        roofline["traffic"] = kops.get("dram_bytes", 0.0) * scale_n
    if top in algo and (nfaces is not None):
        ab = float(algo[top]())
        roofline.update({"bound": "hbm", "unit": "GB/s", "peak": hbm_peak, "peak_source": hbm_src,
                         "algorithmic_bytes_per_step": ab, "achieved": ab / (top_ms * 1e-3) / 1e9})
        roofline["frac"] = roofline["achieved"] / hbm_peak
    else:
        # FP64-pipe bound kernels (north_star: "FP64-pipe utilisation against peak for the Riemann/flux kernel"):
        # flops = FP64 operations of one launch of this kernel on this workload counted by ncu (DFMA = 2)
        fl = (2.0 * kops["dfma"] + kops["dmul"] + kops["dadd"]) * scale_n if kops else None
        roofline.update({"bound": "fp64", "unit": "TFLOP/s", "peak": fp64_peak,
                         "peak_source": "of measured (DFMA microbenchmark run live, mlh_measure_fp64_peak)",
                         "flop_per_step": fl, "achieved": fl / (top_ms * 1e-3) / 1e12 if fl else None})
        roofline["frac"] = roofline["achieved"] / fp64_peak if fl else None
    kernel_rooflines = {}
    for k, ms_k in per_launch.items():
        t = ms_k * 1e-3
        e = {}
        if k in ops:
            o = ops[k]
            e["fp64_frac"] = (2.0 * o["dfma"] + o["dmul"] + o["dadd"]) * scale_n / t / 1e12 / fp64_peak
            e["dram_frac"] = o.get("dram_bytes", 0.0) * scale_n / t / 1e9 / hbm_peak
        if k in algo and nfaces is not None:
            e["hbm_algorithmic_frac"] = float(algo[k]()) / t / 1e9 / hbm_peak
        if k in staged and nfaces is not None and nslots is not None:
            e["staged_bytes_frac"] = float(staged[k]()) / t / 1e9 / hbm_peak
        if not e:
            continue
        hb = max(e.get("dram_frac", 0.0), e.get("staged_bytes_frac", 0.0), e.get("hbm_algorithmic_frac", 0.0))
        fb = e.get("fp64_frac", 0.0)
        e["bound"] = "hbm" if hb >= fb else "fp64"
        e["frac"] = max(hb, fb)
        kernel_rooflines[k] = e
    hbm = {"unit": "GB/s", "peak": hbm_peak, "peak_source": hbm_src,
           "step_achieved": ALL_ALGO_BYTES[D] * n_local / (step_ms_prof * 1e-3) / 1e9}
    hbm["step_frac"] = hbm["step_achieved"] / hbm_peak
    kernels = {k: {"ms_per_step": v[0] / psteps, "launches_per_step": v[1] / psteps} for k, v in prof.items() if v[1]}
    return roofline, kernel_rooflines, hbm, kernels


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="N=8: do not append the Sedov 256^3 run")
    ap.add_argument("--min-seconds", type=float, default=0.6, help="length of the timed region (0: one block of --steps; profiler runs)")
    ap.add_argument("--no-check", action="store_true", help="N>1: skip the bitwise sharded-vs-single check")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    wl = workloads()
    wname = args.workload or "kh2000"
    factory, preset, wdesc = wl[wname]

    # ---------------- reference arm ----------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        side = SIDES[wname]
        if world > 1 and wname == "sedov61":
            side = int(round(61 * world ** (1.0 / 3.0)))
        res = time_reference(factory, preset, side, max(1, args.steps), max(0, min(args.warmup, 1)), budget_s=150.0)
        line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": res["ms_per_step"],
                "higher_is_better": True, "scaling": "weak" if wname == "sedov61" else "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": wdesc, "n_particles_sample": res["n_particles"],
                           "note": "the reference is serial and O(N) per step (plus its O(N N_ghost) ghost search): a per-particle "
                                   "rate measured on a bounded sample of the same IC generator"},
                "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ---------------- B200 arm ----------------
    # A rank that fails must not leave the others waiting in an NCCL call until the launcher's timeout, and no run may
    # sit for more than 25 minutes whatever happens.  The watchdog is a THREAD: a Python signal handler cannot run
    # while the main thread is inside a C call (a blocked ncclAllReduce / cudaStreamSynchronize), a thread can.
    # Once the main workload has been measured its JSON line is parked in `pending`; if one of the optional sections
    # after it (the Sedov 256^3 run at N=8, the sharded-vs-single check) hangs or fails, rank 0 still prints that line,
    # with the section's error recorded in it.
    import threading
    import time
    import traceback

    json_fd = os.dup(1)
    os.dup2(2, 1)  # exactly ONE line may reach stdout; libraries (NCCL prints its version banner there) go to stderr
    state = {"deadline": time.time() + 1500.0, "pending": None, "section": "main", "done": False}
    lock = threading.Lock()

    def _emit_and_exit(why, code):
        with lock:
            if state["done"]:
                return
            state["done"] = True
            line = state["pending"]
        sys.stderr.write("bench.py: rank %d leaves in section '%s': %s\n" % (rank, state["section"], why))
        if line is not None:
            if rank == 0:
                line[state["section"]] = {"error": why}
                os.write(json_fd, (json.dumps(line) + "\n").encode())
            os._exit(0)
        os._exit(code)

    def _watch():
        while True:
            time.sleep(1.0)
            if state["done"]:
                return
            if time.time() > state["deadline"]:
                _emit_and_exit("watchdog: no progress within the section's time limit", 3)
    threading.Thread(target=_watch, daemon=True).start()

    def _enter(section, limit):
        """start an optional section; MLH_BENCH_FAULT=section:raise|hang:rank and MLH_BENCH_SECTION_LIMIT=seconds exist
        so that the way out of a one-rank failure can be exercised on real GPUs (tools/visit_mgpu_final.sh)"""
        state["section"] = section
        state["deadline"] = time.time() + float(os.environ.get("MLH_BENCH_SECTION_LIMIT", limit))
        fault = os.environ.get("MLH_BENCH_FAULT", "").split(":")
        if len(fault) == 3 and fault[0] == section and int(fault[2]) == rank:
            if fault[1] == "raise":
                raise RuntimeError("injected fault in section " + section)
            time.sleep(1e6)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the MFV path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        res = run_workload(args, wname, dist, rank, world, local_rank, want_profile=True, want_e2e=not args.no_e2e)
    except Exception:
        traceback.print_exc()
        _emit_and_exit("exception in the main workload", 2)
    res["world"] = world
    D, n_local = res["D"], res["n_local"]
    roofline, kernel_rooflines, hbm, kernels = rooflines(res, local_rank) if rank == 0 else (None, None, None, None)
    line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": res["scaling"], "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": res["wdesc"], "n_particles": res["n_total"], "dim": D, "mean_neighbours": res["noi_mean"],
                       "kernel_size": res["h"], "abs_mode": "INT_TRUNC (g++/libstdc++ build of the reference)",
                       "l2": "no flush between steps; per-step working set %.0f MB vs 126 MB L2"
                             % (n_local * (ALL_ALGO_BYTES[D] + 4 * (res["noi_mean"] or 32) * 4) / 1e6),
                       "parallelism": "slab%d" % world if world > 1 else "single"},
            "timed_region": res["timed_region"], "clocks": res["clocks"], "e2e": res.get("e2e"),
            "gpu_launches": res["launches"], "gpu_launches_per_step": res["launches"] / float(res["timed_region"]["steps_total"]),
            "roofline": roofline, "hbm_view": hbm, "kernel_rooflines": kernel_rooflines, "num_faces": res["nfaces"],
            "kernels": kernels, "cpu_baseline": None, "device_flags": res["flags"], "also": None, "mgpu_check": None}
    with lock:
        state["pending"] = line

    if world == 8 and args.workload is None and not args.no_also:
        # BASELINE configs[4]: Sedov 3D 256^3 = 16.7 M particles, halo exchange across 8 x B200
        try:
            _enter("also", 420.0)
            a = run_workload(args, "sedov256", dist, rank, world, local_rank, want_profile=False, want_e2e=False)
            line["also"] = {"sedov256": {"workload": a["wdesc"], "n_particles": a["n_total"], "value": a["value"], "unit": UNIT,
                                         "ms_per_step": a["ms_per_step"], "timed_region": a["timed_region"], "n_gpus": world,
                                         "device_flags": a["flags"]}}
        except Exception:
            traceback.print_exc()
            _emit_and_exit("exception", 0)
    if world > 1 and not args.no_check:
        try:
            _enter("mgpu_check", 300.0)
            line["mgpu_check"] = mgpu_bitwise_check(dist, local_rank, rank, world)
        except Exception:
            traceback.print_exc()
            _emit_and_exit("exception", 0)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        _enter("cpu_baseline", 400.0)
        cpu = time_reference(factory, preset, CPU_SIDES[wname], 2, 1, budget_s=25.0)
        line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    with lock:
        state["done"] = True
    if rank == 0:
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        state["done"] = False
        state["pending"], state["section"], state["deadline"] = None, "shutdown", time.time() + 60.0
        try:
            dist.destroy_process_group()
        except Exception:
            pass
        state["done"] = True
    return 0


if __name__ == "__main__":
    sys.exit(main())
