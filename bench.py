#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the pressure-projection hot path.

    python bench.py --gpus N --steps K --warmup W            our CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    the reference's own CPU implementation, measured (never modelled)

A "step" is ONE project() call — fractions, coefficient assembly, multigrid hierarchy, MG-preconditioned CG to the
reference's default tolerance (Residual=1e-4 relative inf-norm), pressure scatter, velocity update. The default workload is
the north-star target, BASELINE.json configs[2]: macliquid3 dam-break with a solid obstacle on 512^3 (level-set free surface,
solid volume fractions, 18.0 M unknowns in 13 % of the box). `value` is grid cells projected per second with the inputs
resident in HBM (restored from pristine device copies inside the timed region); `e2e` is the same through the host-buffer
C-ABI call a Shiokaze module makes (pinned host buffers, H2D and D2H inside the timed region). At N=1 the line also carries
`sub_records.smoke_plume_256` (configs[1], the all-fluid Neumann box the round-1 line was quoted on).

N > 1 (weak scaling, the default): every rank owns one 512^3 copy of the scene, stacked in z into a 512x512x(512 N) box.
`--scaling strong` cuts the n^3 grid itself into N z-slabs (configs[3] flip_splash 512^3 over 1/2/4 GPUs, configs[4] liquid_box
1024^3 over 2/4/8 GPUs); the default N=8 / N=2,4 lines also carry those as `strong` sub-records. Before timing, every multi-GPU
run projects two small grids both ways — cut into N slabs and whole on one GPU — and reports the agreement as `parity_vs_1gpu`.

The reference arm runs the UNMODIFIED reference build (oracle/_ref; the oracle port only where that was never built) through
its own module loader on the SAME scene family at the largest size whose full solve fits the time budget (128^3: its solve
phase is single-threaded and its iteration count grows with N, so 512^3 would take hours). Every number it prints was
measured in that run; nothing is scaled. `config.grid` states the grid it ran.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "projection_throughput"
UNIT = "Mcells/s"
REFERENCE_N = 128        # grid of the reference arm / cpu_baseline leg: the largest full solve of the scene family that takes seconds
WORKLOAD_TEXT = {
    "dambreak_solid": "macliquid3 dam-break + solid obstacle, level-set free surface and solid fractions (BASELINE configs[2])",
    "dambreak": "macliquid3 dam-break, level-set free surface (BASELINE configs[0] geometry)",
    "smoke_plume": "macsmoke3 buoyant plume, all-fluid Neumann box (BASELINE configs[1])",
    "flip_splash": "FLIP liquid splash, rasterised noisy velocity, shell container (BASELINE configs[3])",
    "liquid_box": "synthetic half-filled liquid box, hash-noise velocity (BASELINE configs[4])",
}


# ------------------------------------------------------------------------------------------------------
def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.rows.append(parts)

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
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except ValueError:
                continue
            for name, flag in zip(self.NAMES, r[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
def algorithmic_bytes_per_launch(kernel: str, n_rows: float, precision: str, precond: str = "mg", mid_level=None, tail_level=None):
    """Algorithmic (minimum necessary) bytes one launch of `kernel` moves, counted on unknown rows (DESIGN.md sections 3-4).
    V = CG vector bytes, C = operator coefficient bytes, multigrid is fp32; level l holds n_rows / 8^l rows."""
    V = 4 if precision == "fp32" else 8
    Cc = 8 if precision == "fp64" else 4
    mg = precond == "mg"
    b0 = 4 if (mg and V == 8) else 0          # float copy of r handed to multigrid
    if kernel in ("vcycle_mid", "vcycle_tail"):
        # every level from the launch's first one down: V(2,2) = five passes per level (24 + 28 + 28.5 + 28 sweep bytes, 24.5 residual + restriction), x 8/7
        first = mid_level if kernel == "vcycle_mid" else tail_level
        return None if first is None or first < 0 else n_rows / (8 ** first) * (24 + 28 + 28.5 + 28 + 24.5) * 8.0 / 7.0
    name, _, lvl = kernel.partition("@")
    variant = lvl.lstrip("0123456789g")       # sweep variants: z (x_old = 0, not read), p (+ coarse correction read), d (+ z.r reduced)
    lvl = lvl[:len(lvl) - len(variant)] if variant else lvl
    if lvl and not lvl.isdigit():
        return None                           # gathered coarse levels of a z-slab run ("@g0", ...): tiny, not modelled
    n = n_rows / (8 ** int(lvl)) if lvl else n_rows
    # one launch = one FULL red-black sweep: 4 coefficient arrays + b + x_old read, x_new written (fp32);
    # the first pre-sweep does not read x_old, the first post-sweep also reads the coarse correction (1/8 value per cell)
    sweep_bytes = 24.0 if "z" in variant else (28.5 if "p" in variant else 28.0)
    per_row = {
        "cg_init": 4 * V + b0,                # read b ; write x r s (+ b0)
        "spmv_dot": 2 * V + 4 * Cc,           # read s, wx wy wz dd ; write q (s.q fused)
        "xpay_spmv_dot": (3 * V + 4 + 4 * Cc) if mg else (4 * V + 4 * Cc),   # read z s, wx wy wz dd ; write s q (s = z + beta s fused into the product)
        "axpy2_norm": 6 * V + b0,             # read s q x r ; write x r (+ b0) (norms fused)
        "xpay": (2 * V + 4) if mg else 3 * V, # read z s ; write s
        "dot_rr": V,
        "dot_zb": 8,
        "sweep": sweep_bytes,
        "residual_restrict": 4 * 4 + 4 + 4 + 0.5,   # coefficients, b, x ; coarse b written
        "prolong_add": 8 + 0.5,
    }.get(name)
    return None if per_row is None else per_row * n


def base_tag(k):
    """"sweep@0p" -> "sweep@0": the variant letters of a profiler tag dropped."""
    name, _, lvl = k.partition("@")
    return name + ("@" + lvl.rstrip("zpd") if lvl else "")


def group_kernels(table, ab):
    """Profiler table {tag: (launches, total ms)} -> {kernel function: launches, ms, algorithmic bytes, variants}. The dominant kernel is the
    kernel FUNCTION with the largest summed time: the sweep variants of one level are template instances of the same k_sweep_tma and count
    together, each launch with the algorithmic bytes of its own variant."""
    groups = {}
    for k, (c, t) in table.items():
        if ab(k):
            g = groups.setdefault(base_tag(k), {"launches": 0, "ms": 0.0, "bytes": 0.0, "variants": {}})
            g["launches"] += c; g["ms"] += t; g["bytes"] += ab(k) * c
            g["variants"][k] = {"launches": c, "avg_launch_ms": t / c, "algorithmic_bytes_per_launch": ab(k), "achieved": ab(k) / (t / c * 1e-3) / 1e9}
    return groups


def kernel_source_hash():
    """Identity of the projection's kernel sources: profiles/traffic.json (ncu DRAM bytes per unknown and launch) is only quoted while it was
    captured from THESE sources — a stale capture reads as null, never as a number. (advect.cu is a translation unit of its own and holds none of the
    kernels the table is about.)"""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "shiokaze_b200", "csrc")
    for fn in sorted(os.listdir(d)):
        if fn.endswith((".cuh", ".cu", ".h")) and fn != "advect.cu":
            with open(os.path.join(d, fn), "rb") as f:
                h.update(fn.encode()); h.update(f.read())
    return h.hexdigest()[:16]


def ncu_traffic(groups, dom, rank_rows):
    """DRAM bytes per launch of the dominant kernel from the ncu --set full capture of this round (tools/ncu_traffic.py writes the table)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
    except Exception:
        return None, "no profiles/traffic.json"
    meta = tj.get("_meta", {})
    if meta.get("kernel_source_hash") != kernel_source_hash():
        return None, f"profiles/traffic.json is from other kernel sources (set {meta.get('set')}): not quoted"
    acc, cnt = 0.0, 0
    for k, v in groups[dom]["variants"].items():
        ent = tj.get(k) or tj.get(base_tag(k))
        if not ent:
            return None, f"no ncu capture of {k} in set {meta.get('set')}"
        lvl = base_tag(k).partition("@")[2]
        acc += ent["bytes_per_row"] * rank_rows / (8 ** int(lvl or 0)) * v["launches"]
        cnt += v["launches"]
    return acc / cnt, f"ncu --set full, set {meta.get('set')} ({meta.get('workload')}), dram__bytes_read.sum + dram__bytes_write.sum per unknown x this run's unknowns"


def build_scene(workload, n, zrange=None):
    from shiokaze_b200 import scenes
    return scenes.BENCH_SCENES[workload](n, zrange=zrange) if zrange else scenes.BENCH_SCENES[workload](n)


def workload_string(workload, n, mode, residual):
    return (f"{workload} {n}^3 {mode} ({WORKLOAD_TEXT.get(workload, workload)}), one project() call: fractions + assembly + solve to "
            f"Residual={residual:g} + velocity update")


# ------------------------------------------------------------------------------------------------------
_REF_SCENES = {}


def reference_run(workload, threads, residual):
    """ONE full project() of the reference's CPU implementation on the `workload` scene at REFERENCE_N^3: the unmodified reference build
    through its own module loader (oracle/_ref/f32/ref_driver), or — only where that was never built — the dense C port of the same
    algorithm. Returns measured quantities only."""
    from oracle import refio
    n = REFERENCE_N
    sc = _REF_SCENES.get((workload, n))
    if sc is None:
        sc = _REF_SCENES[(workload, n)] = build_scene(workload, n)
    kind = "reference" if refio.ref_available("f32") else "port"
    t0 = time.perf_counter()
    if kind == "reference":
        r = refio.run_reference(sc, "f32", flags={"Residual": residual}, threads=threads)
        ms, iters, phases, cores = r.ms_project, r.iterations, r.phase_ms, threads
    else:
        from oracle import dense_oracle
        t = time.perf_counter()
        o = dense_oracle.project(sc, residual=residual)
        ms, iters, phases, cores = (time.perf_counter() - t) * 1e3, o.iterations, {}, 1
    return {"ms": ms, "iterations": iters, "phase_ms": phases, "kind": kind, "cores": cores, "n": n, "wall_s": time.perf_counter() - t0,
            "cells": float(n) ** 3}


def reference_sample_text(workload, r, target_n):
    src = "the unmodified reference build (oracle/_ref, Projection=macpressuresolver3 LinSolver=pcg)" if r["kind"] == "reference" else "the dense C port (oracle/dense_oracle.c)"
    ph = ", ".join(f"{k} {v:.0f}" for k, v in r["phase_ms"].items())
    return (f"{workload} {r['n']}^3 — the same scene family as the {target_n}^3 workload at the largest size whose full solve takes seconds — one complete project() "
            f"through {src}: {r['ms']:.0f} ms, {r['iterations']} CG iterations (phases, ms: {ph}); the solve phase is single-threaded in the reference and its "
            f"iteration count grows with N, so its throughput at {target_n}^3 is lower than this figure, which is measured, not extrapolated")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    for _ in range(args.warmup):
        reference_run(args.workload, threads, args.residual)
    t0 = time.perf_counter()
    runs = [reference_run(args.workload, threads, args.residual) for _ in range(args.steps)]
    timed_s = time.perf_counter() - t0
    ms = float(np.mean([r["ms"] for r in runs]))
    value = runs[0]["cells"] / (ms * 1e-3) / 1e6
    n = runs[0]["n"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(args.workload, n, "on the host CPU", args.residual), "grid": [n, n, n], "sample_of": [args.n] * 3,
                   "precision": "fp64 matrix and vectors (FLOAT_TYPE=double), Real=float grids", "precond": "pcg (MIC(0), numerically plain CG: pcg_solver.h:383)",
                   "residual": args.residual, "where": "cpu", "threads": threads,
                   "note": f"the reference has no GPU and no multi-GPU path; every step is one full reference project() at {n}^3, nothing is scaled to {args.n}^3"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": runs[0]["cores"], "kind": runs[0]["kind"], "sample": reference_sample_text(args.workload, runs[0], args.n),
                         "iterations": runs[0]["iterations"], "phase_ms": runs[0]["phase_ms"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "timed_region_s": timed_s,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------------
class Harness:
    """One slab (or whole-grid) solver with pristine device copies of its inputs; PyTorch is only the allocator."""

    def __init__(self, torch, dev, sc, nzg, zr, flags, local, connect=None):
        from shiokaze_b200 import MacPressureSolver3
        self.torch, self.sc, self.dev = torch, sc, dev
        self.S = MacPressureSolver3((sc.nx, sc.ny, nzg), sc.dx, device=local, zrange=zr, **flags)
        if connect:
            connect(self.S)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        self.vel0 = [t(v) for v in sc.vel]
        self.act0 = [t(a) for a in sc.vel_active]
        self.fluid = t(sc.fluid)
        self.solid = t(sc.solid) if sc.solid is not None else None
        self.vel = [torch.empty_like(v) for v in self.vel0]
        self.act = [torch.empty_like(a) for a in self.act0]
        self.pres = torch.zeros(sc.fluid.shape, dtype=torch.float32, device=dev)
        self.pact = torch.zeros(sc.fluid.shape, dtype=torch.uint8, device=dev)

    def step(self):
        for d in range(3):
            self.vel[d].copy_(self.vel0[d]); self.act[d].copy_(self.act0[d])
        return self.S.project_device(self.sc.dt, self.vel, self.act, self.solid, self.fluid, self.sc.fluid_levelset, self.pres, self.pact)

    def host_io_bytes(self):
        sc = self.sc
        h2d = sum(v.nbytes for v in sc.vel) + sum(a.nbytes for a in sc.vel_active) + sc.fluid.nbytes + (sc.solid.nbytes if sc.solid is not None else 0)
        d2h = sum(v.nbytes for v in sc.vel) + sum(a.nbytes for a in sc.vel_active) + sc.fluid.nbytes + sc.fluid.size
        return h2d, d2h

    def close(self):
        self.S.close()


def timed_steps(torch, dist, H, steps, warmup, world, dev, sampler=None):
    """W untimed + K timed project() calls; barrier + synchronise on both sides, CUDA events, max over ranks."""
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    res = None
    for _ in range(max(warmup, 0)):
        res = H.step()
    if sampler:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches, iters = 0, []
    e0.record()
    for _ in range(steps):
        res = H.step()
        launches += res.stats["kernel_launches"] + 6  # + the six restore copies
        iters.append(res.iterations)
    e1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    return ms_total / steps, res, launches, iters, clocks


def e2e_steps(torch, dist, H, steps, warmup, world):
    """The same step through shkz_b200_project_host: pinned host buffers in, pinned host buffers out, wall clock around the call
    (the call returns after its last device-to-host copy), max over ranks."""
    sc = H.sc

    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.from_numpy(np.zeros(1, a.dtype)).dtype, pin_memory=True)
        t.numpy()[...] = a
        return t
    hv0 = [np.ascontiguousarray(v) for v in sc.vel]
    ha0 = [np.ascontiguousarray(a) for a in sc.vel_active]
    hv, ha = [pinned(v) for v in hv0], [pinned(a) for a in ha0]
    hfluid = pinned(sc.fluid)
    hsolid = pinned(sc.solid) if sc.solid is not None else None
    hpres = pinned(np.zeros(sc.fluid.shape, dtype=np.float32))
    hpact = pinned(np.zeros(sc.fluid.shape, dtype=np.uint8))
    total, res = 0.0, None
    w = min(2, warmup)
    for it in range(w + steps):
        for d in range(3):
            hv[d].numpy()[...] = hv0[d]; ha[d].numpy()[...] = ha0[d]
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        _, _, res = H.S.project(sc.dt, [t.numpy() for t in hv], [t.numpy() for t in ha], hsolid.numpy() if hsolid is not None else None,
                                hfluid.numpy(), sc.fluid_levelset, pressure_out=hpres.numpy(), pressure_active_out=hpact.numpy())
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=H.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        if it >= w:
            total += dt
    return total / steps, res


def profile_one_step(H):
    H.S.profile(True)
    res = H.step()
    table = H.S.profile_table()
    H.S.profile(False)
    return table, res


def roofline_of(table, res, rank_rows, precision, precond, ms_solve):
    """table: per-kernel CUDA-event times of ONE profiled project(); ms_solve: solve phase of an UNprofiled step (events around every launch slow the step down)."""
    peak, how = measured_peak()
    ab = lambda k: algorithmic_bytes_per_launch(k, rank_rows, precision, precond, res.stats.get("mg_mid_level"), res.stats.get("mg_tail_level"))
    groups = group_kernels(table, ab)
    if not groups:
        return None
    dom = max(groups, key=lambda k: groups[k]["ms"])
    cnt, tot = groups[dom]["launches"], groups[dom]["ms"]
    per_launch = groups[dom]["bytes"] / cnt
    achieved = per_launch / (tot / cnt * 1e-3) / 1e9
    traffic, traffic_source = ncu_traffic(groups, dom, rank_rows)
    total_profiled = sum(v[1] for v in table.values())
    alg_total = sum((ab(k) or 0) * v[0] for k, v in table.items())
    return {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_source": traffic_source, "peak_source": how, "launches": cnt, "avg_launch_ms": tot / cnt, "algorithmic_bytes_per_launch": per_launch,
            "share_of_step": tot / total_profiled if total_profiled else None, "variants": groups[dom]["variants"],
            "solve_whole": {"algorithmic_GB": alg_total / 1e9, "ms": ms_solve, "achieved": alg_total / 1e9 / (ms_solve * 1e-3),
                            "frac": alg_total / 1e9 / (ms_solve * 1e-3) / peak,
                            "what": "every solve kernel of one project(): algorithmic bytes (counted on unknown rows) / CUDA-event time of the solve phase / peak"},
            "active_tiles": [res.stats.get("active_tiles"), res.stats.get("total_tiles")],
            "by_kernel_ms": {k: round(v[1], 4) for k, v in sorted(table.items(), key=lambda kv: -kv[1][1])[:16]}}


# ---- the step before the projection (first part): advect_vector on the bench scene -----------------------------------------------
# Algorithmic bytes per ACTIVE face of one MacCormack advect_vector (csrc/advect.cu), Real = float: forward kernel reads u (4) + mask (1) and, per
# cell, the level set (4/3 per face), writes the forward value (4) and the limiter record min / max / narrow-band flag (4 + 4 + 1); backward + limiter kernel
# reads the forward value (4), mask (1), the record (9) and u (4), writes u (4). Stencil neighbours are re-reads of the same arrays (cache hits) and get no credit.
ADVECT_BYTES_PER_ACTIVE_FACE = (4 + 1 + 4.0 / 3.0 + 4 + 9) + (4 + 1 + 9 + 4 + 4)


def advect_dt(sc, cells=2.0):
    """Time step that carries the scene's fastest face over `cells` cells (the simulators run at CFL 1-3; the projection scenes' dt moves a dam-break by 0.35)."""
    vmax = max(float(np.abs(v).max()) for v in sc.vel)
    return cells * sc.dx / vmax if vmax > 0 else sc.dt


def reference_advect(workload, threads):
    """ONE advect_vector of the unmodified reference module (oracle/_ref, Advection=macadvection3) on the workload's scene at REFERENCE_N^3, all host threads."""
    from oracle import refio
    import dataclasses
    if not refio.ref_available("f32"):
        return None
    n = REFERENCE_N
    sc = _REF_SCENES.get((workload, n))
    if sc is None:
        sc = _REF_SCENES[(workload, n)] = build_scene(workload, n)
    r = refio.run_reference(dataclasses.replace(sc, dt=advect_dt(sc)), "f32", threads=threads, advect="vector")
    faces = float(sum(int(a.sum()) for a in sc.vel_active))
    return {"ms": r.ms_project, "active_faces": faces, "n": n, "cores": threads}


def advect_sub_record(torch, dev, local, sc, workload, n, steps, with_cpu):
    from shiokaze_b200 import MacAdvection3
    A = MacAdvection3(sc.shape, sc.dx, device=local)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    u0, act, fluid = [t(v) for v in sc.vel], [t(a) for a in sc.vel_active], t(sc.fluid)
    u = [torch.empty_like(v) for v in u0]
    dt = advect_dt(sc)
    faces = float(sum(int(a.sum()) for a in sc.vel_active))

    def step():
        for d in range(3):
            u[d].copy_(u0[d])
        return A.advect_vector_device([x.data_ptr() for x in u], [x.data_ptr() for x in act], fluid.data_ptr(), dt)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms_kernels, launches = 0.0, 0
    e0.record()
    for _ in range(steps):
        st = step()
        ms_kernels += st["ms_advect"]
        launches += st["kernel_launches"] + 3
    e1.record()
    torch.cuda.synchronize()
    ms_step = e0.elapsed_time(e1) / steps
    ms_kernels /= steps
    # end to end: the call the Shiokaze module makes, page-locked host grids in and out
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    hu, ha, hf = [pin(v) for v in sc.vel], [pin(a) for a in sc.vel_active], pin(sc.fluid)
    import ctypes as C
    from shiokaze_b200 import capi
    e2e = []
    for it in range(2 + 3):
        for d in range(3):
            hu[d].copy_(torch.from_numpy(sc.vel[d]))
        stt = capi.AdvectStats()
        t0 = time.perf_counter()
        capi.check_advect(capi.lib().shkz_b200_advect_vector_host(A._h, dt, (C.c_void_p * 3)(*[x.data_ptr() for x in hu]), (C.c_void_p * 3)(*[x.data_ptr() for x in ha]),
                                                                 hf.data_ptr(), C.byref(A.params), C.byref(stt)))
        if it >= 2:
            e2e.append(time.perf_counter() - t0)
    A.close()
    peak, peak_src = measured_peak()
    alg = ADVECT_BYTES_PER_ACTIVE_FACE * faces
    rec = {"config": f"advect_vector (MacCormack, trilinear: the reference's defaults) of the {workload} {n}^3 velocity, dt = 2 cells of the fastest face; "
                     f"{int(faces)} active faces of {sum(v.size for v in sc.vel)}",
           "ms_per_step": ms_step, "value": faces / (ms_step * 1e-3) / 1e6, "unit": "Mfaces/s (active faces)", "gpu_launches_per_step": launches // steps,
           "e2e_ms_per_step": float(np.mean(e2e)) * 1e3, "e2e_h2d_bytes": int(stt.h2d_bytes), "e2e_d2h_bytes": int(stt.d2h_bytes),
           "e2e_host_copies": "sparse" if stt.host_copies else "dense",
           "roofline": {"bound": "hbm", "kernel": "k_advect_faces (forward + record, backward + limiter)", "achieved": alg / (ms_kernels * 1e-3) / 1e9, "peak": peak,
                        "unit": "GB/s", "frac": alg / (ms_kernels * 1e-3) / 1e9 / peak, "peak_source": peak_src, "kernels_ms": ms_kernels,
                        "algorithmic_bytes_per_active_face": ADVECT_BYTES_PER_ACTIVE_FACE, "traffic": None}}
    if with_cpu:
        r = reference_advect(workload, os.cpu_count() or 1)
        if r:
            rec["cpu_baseline"] = {"value": r["active_faces"] / (r["ms"] * 1e-3) / 1e6, "unit": "Mfaces/s (active faces)", "cores": r["cores"], "kind": "reference",
                                   "sample": f"{workload} {r['n']}^3, one advect_vector of the unmodified reference module (oracle/_ref, Advection=macadvection3): {r['ms']:.0f} ms, measured"}
    return rec


def solve_record(res, n_rows, iters):
    phase = {k: res.stats[k] for k in ("ms_assemble", "ms_setup", "ms_solve", "ms_update")}
    return {"iterations": iters[-1], "reresid": res.reresid, "converged": res.converged, "n_rows": int(n_rows),
            "cell_iters_per_s": n_rows * iters[-1] / (phase["ms_solve"] * 1e-3) if phase["ms_solve"] > 0 else None, **phase}


def parity_vs_1gpu(torch, dist, sdist, rank, world, local):
    """Before timing a multi-GPU run: two small global grids projected cut into `world` slabs AND whole on this rank's own GPU (fp64, Residual=1e-10).
    Every rank compares its slab of the outputs; the line carries the worst rank. Masks must be equal, velocity within 1e-6 rel. L2."""
    from shiokaze_b200 import MacPressureSolver3, scenes
    out = {}
    ok_all = True
    for name, make in (("dambreak_solid_64", lambda: scenes.dambreak(64, True)), ("smoke_plume_64", lambda: scenes.smoke_plume(64))):
        sc = make()
        flags = dict(Precision="fp64", Precond="mg", Residual=1e-10)
        W = MacPressureSolver3((sc.nx, sc.ny, sc.nz), sc.dx, device=local, **flags)
        whole = W.project_scene(sc)
        W.close()
        k0, k1 = sdist.slab_range(sc.nz, rank, world)
        part = sdist.split_dense(sc, world)[rank]
        S = MacPressureSolver3((sc.nx, sc.ny, sc.nz), sc.dx, device=local, zrange=(k0, k1), **flags)
        sdist.connect(S, rank, world)
        mine = S.project_scene(part)
        S.close()
        zs = [slice(k0, k1), slice(k0, k1), slice(k0, k1 + 1)]
        masks = all(np.array_equal(mine["vel_active"][d], whole["vel_active"][d][zs[d]]) for d in range(3)) and \
            np.array_equal(mine["pressure_active"], whole["pressure_active"][k0:k1])
        num = sum(float(((mine["vel"][d].astype(np.float64) - whole["vel"][d][zs[d]]) ** 2).sum()) for d in range(3))
        den = sum(float((whole["vel"][d].astype(np.float64) ** 2).sum()) for d in range(3))
        t = torch.tensor([num, 0.0 if masks else 1.0], dtype=torch.float64, device=torch.device("cuda", local))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        rel = (float(t[0].item()) / den) ** 0.5 if den > 0 else 0.0
        masks_all = float(t[1].item()) == 0.0
        ok = masks_all and rel < 1e-6
        ok_all = ok_all and ok
        out[name] = {"grid": [sc.nx, sc.ny, sc.nz], "slabs": world, "masks_equal": masks_all, "vel_rel_l2": rel,
                     "iterations": [mine["result"].iterations, whole["result"].iterations], "ok": ok}
    out["ok"] = ok_all
    out["bar"] = "activity masks equal, velocity <= 1e-6 rel. L2 between the slab run and the whole-grid run on one GPU (fp64, Residual=1e-10)"
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from shiokaze_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    if capi.lib().shkz_b200_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: libshkz_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    sdist = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        from shiokaze_b200 import dist as sdist

    flags = dict(Precision=args.precision, Precond=args.precond, Residual=args.residual, MGPreSweeps=args.pre, MGPostSweeps=args.post,
                 CheckEvery=args.check_every)

    def harness(workload, n, strong):
        nzg = n if strong else n * world
        if strong and n % world:
            raise SystemExit(f"--scaling strong: n={n} is not divisible by {world} slabs")
        nzl = nzg // world
        zr = (rank * nzl, (rank + 1) * nzl)
        sc = sdist.slab_scene(workload, n, nzg, zr) if world > 1 else build_scene(workload, n)
        return Harness(torch, dev, sc, nzg, zr, flags, local, connect=(lambda S: sdist.connect(S, rank, world)) if world > 1 else None), nzg

    parity = parity_vs_1gpu(torch, dist, sdist, rank, world, local) if world > 1 else None

    n = args.n
    strong = args.scaling == "strong"
    H, nzg = harness(args.workload, n, strong)
    sampler = ClockSampler(local) if rank == 0 else None
    ms_step, res, launches, iters, clocks = timed_steps(torch, dist, H, args.steps, args.warmup, world, dev, sampler)
    cells_global = float(n) * n * nzg
    value = cells_global / (ms_step * 1e-3) / 1e6
    n_rows = res.stats["n_rows_global"] if world > 1 else res.n_rows

    # ---- end to end through the host-buffer C-ABI (what the Shiokaze module calls), pinned host memory, every rank its slab ----
    h2d, d2h = H.host_io_bytes()
    e2e_s, e2e_res = e2e_steps(torch, dist, H, args.steps, args.warmup, world)
    e2e_value = cells_global / e2e_s / 1e6

    # ---- roofline of the dominant kernel: CUDA events around every launch of one extra project() ----
    table, pres = profile_one_step(H)
    roofline = roofline_of(table, pres, res.n_rows / world, args.precision, args.precond, res.stats["ms_solve"]) if rank == 0 else None
    main_scene = H.sc
    H.close()
    del H
    torch.cuda.empty_cache()

    # ---- secondary records ----
    sub = {}
    if world == 1 and not args.no_sub_records:
        try:
            sub[f"advect_vector_{n}"] = advect_sub_record(torch, dev, local, main_scene, args.workload, n, max(5, args.steps // 2), not args.no_cpu_baseline)
        except Exception as e:   # a secondary record never takes the headline line down with it
            sub[f"advect_vector_{n}"] = {"error": repr(e)}
        torch.cuda.empty_cache()
    del main_scene
    if world == 1 and not args.no_sub_records and (args.workload, n) != ("smoke_plume", 256):
        Hs, _ = harness("smoke_plume", 256, False)
        ms_s, res_s, _, it_s, _ = timed_steps(torch, dist, Hs, max(5, args.steps // 2), 3, 1, dev)
        e2e_ss, _ = e2e_steps(torch, dist, Hs, 5, 2, 1)
        tab_s, pres_s = profile_one_step(Hs)
        rf = roofline_of(tab_s, pres_s, res_s.n_rows, args.precision, args.precond, res_s.stats["ms_solve"])
        sub["smoke_plume_256"] = {"config": workload_string("smoke_plume", 256, "on one GPU", args.residual), "ms_per_step": ms_s,
                                  "value": 256.0 ** 3 / (ms_s * 1e-3) / 1e6, "unit": UNIT, "e2e_ms_per_step": e2e_ss * 1e3,
                                  "e2e_value": 256.0 ** 3 / e2e_ss / 1e6,
                                  "solve": solve_record(res_s, res_s.n_rows, it_s),
                                  "roofline": {k: rf[k] for k in ("kernel", "achieved", "peak", "frac", "solve_whole")} if rf else None}
        Hs.close()
        del Hs
        torch.cuda.empty_cache()
    strong_rec = None
    if world > 1 and not strong and not args.no_sub_records:
        # the strong-scaling targets of BASELINE.json ride along: configs[3] (FLIP splash 512^3) on 2 / 4 GPUs, configs[4] (liquid box 1024^3) on 8
        sw, sn = ("liquid_box", 1024) if world == 8 else ("flip_splash", 512)
        Hs, _ = harness(sw, sn, True)
        ms_s, res_s, _, it_s, _ = timed_steps(torch, dist, Hs, 5, 3, world, dev)
        strong_rec = {"config": workload_string(sw, sn, f"cut into {world} z-slabs", args.residual), "grid": [sn, sn, sn], "ms_per_step": ms_s,
                      "value": float(sn) ** 3 / (ms_s * 1e-3) / 1e6, "unit": UNIT, "solve": solve_record(res_s, res_s.stats["n_rows_global"], it_s)}
        Hs.close()
        del Hs

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = reference_run(args.workload, os.cpu_count() or 1, args.residual)
        cpu_baseline = {"value": r["cells"] / (r["ms"] * 1e-3) / 1e6, "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                        "sample": reference_sample_text(args.workload, r, n), "ms_per_solve": r["ms"], "grid": [r["n"]] * 3, "iterations": r["iterations"]}

    if rank == 0:
        mode = "cut into z-slabs" if strong else ("per GPU, stacked in z" if world > 1 else "on one GPU")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": {"mixed": "f64", "fp64": "f64", "fp32": "f32"}[args.precision], "data": "synthetic",
            "config": {"workload": workload_string(args.workload, n, mode, args.residual), "grid": [n, n, nzg], "slab_per_gpu": [n, n, nzg // world],
                       "parallelism": f"z-slab x{world}", "precision": args.precision, "precond": args.precond, "mg_sweeps": [args.pre, args.post],
                       "residual": args.residual,
                       "l2_policy": "inputs and solver working set of one step (GBs at 512^3, > 1 GB at 256^3) exceed the 126 MB L2; no explicit flush",
                       "where": "gpu"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(e2e_res.stats["h2d_bytes"]) * world,
                    "d2h_bytes_per_step": int(e2e_res.stats["d2h_bytes"]) * world, "ms_per_step": e2e_s * 1e3,
                    "ms_h2d": e2e_res.stats["ms_h2d"], "ms_d2h": e2e_res.stats["ms_d2h"],
                    "host_copies": "sparse" if e2e_res.stats["host_copies"] else "dense",
                    "whole_array_bytes_per_step": [h2d * world, d2h * world],
                    "what": "shkz_b200_project_host with page-locked host buffers, wall clock around the call, max over ranks; bytes = what the call moved over PCIe "
                            "as counted by the library (stats.h2d_bytes / d2h_bytes, rank 0 x ranks): on a liquid scene only the level set, the masks, and the "
                            "velocity / solid nodes / pressure around wet cells travel (csrc/kernels_xfer.cuh); whole_array_bytes_per_step = every array whole"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "solve": solve_record(res, n_rows, iters),
        }
        if sub:
            line["sub_records"] = sub
        if strong_rec:
            line["strong"] = strong_rec
        if parity is not None:
            line["parity_vs_1gpu"] = parity
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="dambreak_solid")
    ap.add_argument("--n", "--grid", dest="n", type=int, default=512, help="grid size n (n^3 cells per GPU, or in total with --scaling strong); --grid is the spelling to use under torchrun, whose own parser finds --n ambiguous")
    ap.add_argument("--precision", default="mixed", choices=["mixed", "fp64", "fp32"])
    ap.add_argument("--precond", default="mg", choices=["mg", "none"])
    ap.add_argument("--pre", type=int, default=2)
    ap.add_argument("--post", type=int, default=2)
    ap.add_argument("--residual", type=float, default=1e-4)
    ap.add_argument("--check-every", type=int, default=4)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="N>1: n^3 per GPU (weak) or the n^3 grid cut into N slabs (strong)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub-records", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: at least three warm-up steps
    return run_reference_arm(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
