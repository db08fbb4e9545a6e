#!/usr/bin/env python
"""bench.py — LM iterations/s of the B200-native Levenberg-Marquardt inner loop on a synthetic BAL problem.

A "step" is one LM iteration (damping, matrix-free Schur PCG solve, back-substitution, update, cost, accept or
reject + re-linearisation) of the optimisation trajectory under the reference's BAL protocol
(examples/bal.cu:284-309: lambda0 1e-4, PCG 10 iterations / tolerance 1.0 / rejection ratio 5.0).
W warm-up iterations start from the initial state, then exactly K iterations are timed.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
  python bench.py --impl reference ...                     (CPU arm: the oracle port on the host cores)

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definition of every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: native libraries (NCCL prints its "NCCL version ..." banner on fd 1) are sent to
# stderr for the whole run, and the line is written to the saved descriptor at the end
os.environ.setdefault("NCCL_DEBUG", "WARN")  # never override the caller's setting (the driver reads NCCL's INFO lines)
sys.stdout.flush()
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


from graphite_b200 import synthetic  # noqa: E402

METRIC = "LM iterations/s (synthetic BAL, Schur PCG)"
UNIT = "LM it/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed regions (B200_PROFILING.md clocks line).

    The device-resident timed region lasts tens of milliseconds, far below nvidia-smi's polling period, so the
    samples are taken in-process through NVML (pynvml) every ~2 ms by a thread; `mark(True/False)` brackets the timed
    regions and only samples inside them count.  nvidia-smi is the fallback when NVML cannot be loaded."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, device: int):
        self.device = device
        self.samples = []  # (in_region, sm_mhz, reasons bitmask)
        self.in_region = False
        self.stop_flag = False
        self.thread = None
        self.max_mhz = None
        self.mode = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.device
            if vis:
                try:
                    idx = int(vis.split(",")[self.device])
                except (ValueError, IndexError):
                    idx = self.device
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.mode = "nvml"
        except Exception:
            self.mode = "smi"
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def _loop(self):
        if self.mode == "nvml":
            nv = self.nv
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
            while not self.stop_flag:
                try:
                    mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    rs = int(get_reasons(self.h))
                    self.samples.append((self.in_region, mhz, rs))
                except Exception:
                    pass
                time.sleep(0.002)
        else:
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            bits = [0x8, 0x40, 0x20, 0x4]
            while not self.stop_flag:
                try:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.device)], capture_output=True, text=True, timeout=5).stdout
                    r = [t.strip() for t in out.strip().split(",")]
                    rs = sum(b for b, v in zip(bits, r[2:6]) if v.lower().startswith("active"))
                    self.max_mhz = float(r[1])
                    self.samples.append((self.in_region, float(r[0]), rs))
                except Exception:
                    time.sleep(0.05)

    def mark(self, inside: bool):
        self.in_region = inside

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=10)
        rows = [s for s in self.samples if s[0]] or self.samples
        where = "timed regions" if any(s[0] for s in self.samples) else "whole run (no sample fell inside a timed region)"
        sm = [s[1] for s in rows]
        mask = 0
        for s in rows:
            mask |= s[2]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(n for b, n in self.REASONS.items() if mask & b), "samples": len(sm),
                "source": self.mode, "sampled": where}


GOLDEN_PRECISION = {"f64-f64": "FP64-FP64", "f32-f32": "FP32-FP32", "f64-f32": "FP64-FP32", "f64-bf16": "FP64-BF16"}


def golden_parity(workload, precision, solver, chi2_end, accepted_vector):
    """Compare this run (warm-up + steps, from the initial state, reference protocol) with the committed run of the
    UNMODIFIED reference on a B200 (tests/golden/*.json, generated by oracle/make_golden.py): relative difference of
    chi2 after the same number of LM iterations and the accept / reject decision of every iteration."""
    name = f"{workload}__{solver}__{GOLDEN_PRECISION.get(precision, precision)}.json"
    path = os.path.join(ROOT, "tests", "golden", name)
    if not os.path.exists(path):
        return {"golden": None, "note": f"no reference run committed for {name}"}
    with open(path) as fh:
        g = json.load(fh)
    table = g["table"]
    n = len(accepted_vector)
    if n == 0 or n > len(table):
        return {"golden": name, "note": f"reference run has {len(table)} iterations, this run {n}"}
    ref_end = table[n - 1][2]
    ref_acc = [bool(row[2] < row[1]) for row in table[:n]]
    rel = abs(chi2_end - ref_end) / abs(ref_end)
    tol = 1e-6 if precision == "f64-f64" else 1e-4  # BASELINE.json north_star tolerances
    # reduced precision: the reference itself is not reproducible from run to run (float atomics), so a decision may
    # flip where rho is at rounding level; the cost bound is what the north star states
    dec = [bool(a) for a in accepted_vector] == ref_acc
    ok = rel <= tol and (dec or precision != "f64-f64")
    return {"golden": name, "iterations": n, "chi2": chi2_end, "chi2_reference": ref_end, "rel": rel, "tolerance": tol,
            "decisions_equal": dec, "ok": bool(ok)}


def reference_gpu_leg(prob, precision, solver, warmup, steps):
    """The UNMODIFIED reference's own GPU path (oracle/_ref/ref_bal, compiled from /root/reference for sm_100 by
    oracle/Makefile; test infrastructure) on the same B200, same problem, same protocol: its per-iteration times for LM
    iterations warmup .. warmup+steps-1 (levenberg_marquardt.hpp:212-221 table), set-up time apart."""
    import tempfile
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_bal")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/ref_bal not built"}
    if precision not in GOLDEN_PRECISION:
        return {"unavailable": f"precision {precision}"}
    try:
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "p.gbal")
            synthetic.write_gbal(prob, path)
            t0 = time.perf_counter()
            out = subprocess.run([exe, path, "--solver", solver, "--precision", GOLDEN_PRECISION[precision], "--iterations",
                                  str(warmup + steps), "--lambda", "1e-4"], capture_output=True, text=True, timeout=900)
            wall = time.perf_counter() - t0
        rows = []
        for line in out.stdout.splitlines():
            tok = line.split()
            if len(tok) == 6:
                try:
                    rows.append([int(tok[0])] + [float(v) for v in tok[1:]])
                except ValueError:
                    pass
        sel = [r for r in rows if warmup <= r[0] < warmup + steps]
        if out.returncode != 0 or not sel:
            return {"unavailable": f"ref_bal rc {out.returncode}, {len(rows)} rows: {out.stderr[-200:]}"}
        sec = sum(r[4] for r in sel)
        return {"value": len(sel) / sec, "unit": UNIT, "ms_per_step": 1e3 * sec / len(sel), "iterations": [sel[0][0], sel[-1][0]],
                "setup_seconds": rows[0][5] - rows[0][4], "wall_seconds": wall, "chi2_end": sel[-1][2],
                "what": "sfu-rsl/graphite v0.5.0 GPU path (PCGSchurSolver / PCGSolver), unmodified headers compiled for sm_100, "
                        "same B200, same problem and protocol; host-clock time per iteration as the reference prints it"}
    except Exception as e:  # the leg is informative: never fail the bench on it
        return {"unavailable": repr(e)[:200]}


def partition_points(prob, nranks: int, rank: int):
    """Contiguous point ranges balanced by observation count."""
    from graphite_b200.distributed import partition_by_point
    return partition_by_point(prob, nranks, rank)


def algorithmic_bytes(nc, npts, m, nrows, ntiles, sT, sS):
    """Bytes one launch of the Schur-product kernel has to move in this layout (DESIGN.md "Kernels"):
    Jacobians 24 values + 4 B packed meta per observation, the per-tile segment/point tables, W per point,
    camera-vector rows in and partial rows out per (super-tile, camera)."""
    V = 9 * nc * sT
    product = m * (24 * sS + 4) + ntiles * (2400 - 1024) + npts * 6 * sT + 2 * nrows * 9 * sT
    # SURVEY section 8(d) K4 implicit, whole PCG iteration, for reference next to it
    survey_k4 = 27 * m * sS + 4 * m + 9 * npts * sS + 81 * nc * sS + 10 * V
    return product, survey_k4


def lm_iteration_bytes(nc, npts, m, sT, sS, k_pcg, accepted_fraction):
    """algorithmic bytes of one LM iteration on the implicit-Schur path, with the PCG iterations
    actually executed and the share of iterations that re-linearise (accepted steps)."""
    V = 9 * nc * sT
    k1k2 = m * (2 * sT + 8) + 9 * nc * sT + 3 * npts * sT + 2 * m * sT + 27 * m * sS + 81 * nc * sS + 9 * npts * sS + (9 * nc + 3 * npts) * sT
    k3 = 2 * 9 * npts * sS + 27 * m * sS + (9 * nc + 3 * npts) * sT + V
    k4 = 27 * m * sS + 4 * m + 9 * npts * sS + 81 * nc * sS + 10 * V
    k5 = (27 * m * sS + 4 * m + 9 * npts * sS + 6 * npts * sT + V) + (12 * npts * sT + 4 * V + m * (2 * sT + 8))
    return accepted_fraction * k1k2 + k3 + k_pcg * k4 + k5


def run_reference(args):
    """CPU arm: the oracle port (the reference's Eigen CPU path cannot be built here: Eigen is absent)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.binding import Oracle, default_options
    prob = synthetic.make_named(args.workload)
    cores = os.cpu_count() or 1
    O = Oracle(prob, "f64")
    opts = default_options(iterations=args.warmup + args.steps, threads=cores)
    O.lm_begin(opts)
    ks = []
    for _ in range(args.warmup):
        out, go = O.lm_step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out, go = O.lm_step()
        ks.append(int(out[3]))
    dt = time.perf_counter() - t0
    value = args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(prob, "f64-f64", args.gpus, args.schur_mode), solver=args.solver),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} LM iterations after {args.warmup} warm-up iterations of the same workload "
                                   f"(explicit Schur + PCG, OpenMP); pcg iterations per step {ks}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "final_chi2": float(out[1]),
    }
    emit(line)


def workload_config(prob, precision, n_gpus, schur_mode="auto"):
    nc, npts, m = prob.shape()
    return {"workload": f"{prob.name}: {nc} cams / {npts} pts / {m} obs, seed 0, {precision.upper()}, points eliminated, "
                        f"lambda0 1e-4, PCG 10 it / tol 1.0 / rejection 5.0, diagonal damping, Jacobi scaling",
            "cams": nc, "points": npts, "observations": m, "precision": precision,
            "schur": {"auto": "library rule (implicit at this size)", "implicit": "implicit (matrix-free)",
                      "explicit": "explicit (stored S)"}[schur_mode], "parallelism": f"points partitioned over {n_gpus} GPU(s), cameras replicated",
            "l2": "no flush needed: each pass streams the 0.96 GB Jacobian store (FP64) which is larger than the 126 MB L2"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=45)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="venice-1778")
    ap.add_argument("--precision", default="f64-f64")
    ap.add_argument("--solver", default="pcg-schur", choices=["pcg-schur", "pcg"],
                    help="pcg-schur: PCGSchurSolver (headline); pcg: the reference's full-system PCGSolver (its mixed-precision path)")
    ap.add_argument("--schur-mode", default="auto", choices=["auto", "implicit", "explicit"],
                    help="form of the Schur complement (gb_pcg_options.schur_mode): matrix-free, stored, or the library's rule")
    ap.add_argument("--super-tile-obs", type=int, default=0,
                    help="tuning: target observations per super-tile (gb_problem_desc.super_tile_observations), 0 = library default")
    ap.add_argument("--device-only", action="store_true",
                    help="tuning sweeps: only the device-resident arm (no e2e, CPU or reference legs); not a reportable line")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-gpu", action="store_true",
                    help="skip the leg that runs the unmodified reference's GPU path (oracle/_ref/ref_bal) on the same box")
    ap.add_argument("--cpu-baseline-steps", type=int, default=4)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    from graphite_b200 import binding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    prob = synthetic.make_named(args.workload)
    nc, npts, m = prob.shape()
    local = partition_points(prob, world, rank) if world > 1 else prob
    ctx = binding.Context(local_rank)
    if world > 1:
        uid = [binding.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(world, rank, uid[0])
    tname, sname = args.precision.split("-")
    T = np.float64 if tname == "f64" else np.float32
    sT, sS = (8 if tname == "f64" else 4), {"f64": 8, "f32": 4, "bf16": 2}[sname]
    t_struct = time.perf_counter()
    P = binding.Problem(ctx, local.cam_idx, local.pt_idx, local.n_cams, local.n_pts, args.precision, partition=world > 1,
                        super_tile_observations=args.super_tile_obs)
    structure_seconds = time.perf_counter() - t_struct  # one-time: sort, tiles, segments, camera CSR, uploads
    info = P.info()

    # pinned host copies of the inputs (the e2e arm copies from these every step)
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=T)).pin_memory()
        return t
    h_obs, h_cams, h_pts = pinned(local.obs), pinned(local.cams), pinned(local.pts)
    out_cams, out_pts = torch.empty_like(h_cams).pin_memory(), torch.empty_like(h_pts).pin_memory()
    P.set_observations_raw(h_obs.data_ptr())
    P.set_vertices_raw(h_cams.data_ptr(), h_pts.data_ptr())

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident arm: W warm-up LM iterations, then K timed --------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    traj_w, res_w = P.lm(iterations=args.warmup, solver=args.solver, schur_mode=args.schur_mode)
    barrier()
    sampler.mark(True)
    l0 = ctx.kernel_launches()
    t0 = time.perf_counter()
    traj, res = P.lm(iterations=args.steps, initial_damping=res_w["final_damping"], initial_nu=res_w["final_nu"],
                     resume=True, profile_product=True, solver=args.solver, schur_mode=args.schur_mode)
    barrier()
    sampler.mark(False)
    wall = time.perf_counter() - t0
    launches = ctx.kernel_launches() - l0
    seconds = max_over_ranks(res["seconds_total"])  # CUDA events on the context stream, max over ranks
    steps_done = int(res["iterations"])
    value = steps_done / seconds
    # parity with the committed run of the unmodified reference: same protocol, same number of LM iterations from the
    # same initial state (all ranks hold identical scalars; rank 0 reports)
    full_traj = np.concatenate([traj_w, traj]) if len(traj_w) else traj
    parity = golden_parity(args.workload, args.precision, args.solver, float(res["final_chi2"]),
                           [bool(row[1] < row[0]) for row in full_traj])

    # ---- roofline of the dominant kernel, timed live in the run above ------------------------------------------------
    # k_pcg_solve = one launch per LM iteration = the whole PCG solve (k executed iterations of the matrix-free Schur
    # product + row sums + exchange + vector updates).  Time: CUDA events around each launch on the context stream
    # (gb_lm_result.seconds_pcg, summed over the timed LM iterations).  Algorithmic bytes of a launch: k x the bytes one
    # product has to move in this layout.  The product phase alone (in-kernel globaltimer stamps of CTA 0: phase start ->
    # after the grid barrier that ends it) is listed next to it.
    peak, peak_kind = load_peaks()
    prod_bytes, survey_k4 = algorithmic_bytes(nc, local.n_pts, info["n_obs"], info["n_partial_rows"], info["n_tiles"], sT, sS)
    if args.solver == "pcg-schur":  # (the protocol's 10 PCG iterations: within the 16 the library keeps point sums for)
        prod_bytes += 3 * local.n_pts * sT  # the point sums t_p(p_k) every iteration leaves for the Jacobian-free back-substitution
    k_total = max(int(res["pcg_iterations_total"]), 1)
    pcg_seconds = max(res["seconds_pcg"], 1e-12)
    launches_timed = max(steps_done, 1)
    bytes_per_launch = prod_bytes * k_total / launches_timed
    ms_per_launch = 1e3 * pcg_seconds / launches_timed
    achieved = prod_bytes * k_total / pcg_seconds / 1e9
    n_prod = max(int(res["product_launches"]), 1)
    prod_ms = 1e3 * res["product_seconds"] / n_prod
    prod_gbps = prod_bytes / (prod_ms * 1e-3) / 1e9 if prod_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "k_pcg_solve (persistent cooperative kernel: the whole PCG solve, one launch per LM iteration; "
                                           "per PCG iteration the matrix-free Schur product on the TMA pipeline + row sums + exchange + updates)",
                "achieved": achieved, "peak": peak, "peak_kind": peak_kind + " (MEASURED_PEAKS.json hbm_gbs, burst copy)",
                "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "bytes_per_launch": bytes_per_launch, "launches_timed": launches_timed, "ms_per_launch": ms_per_launch,
                "pcg_iterations_per_launch": k_total / launches_timed, "bytes_per_pcg_iteration": prod_bytes,
                "share_of_step": pcg_seconds / max(res["seconds_total"], 1e-12),
                "product_phase": {"ms": prod_ms, "achieved": prod_gbps, "frac": prod_gbps / peak, "iterations_timed": int(res["product_launches"]),
                                  "how": "globaltimer stamps of CTA 0 inside the kernel: phase start to the exit of the grid barrier that ends the phase"},
                "survey_k4_bytes_per_pcg_iteration": survey_k4}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            with open(prof) as fh:
                tr = json.load(fh)
            key = f"{args.workload}:{args.precision}:{world}:per_pcg_iteration"
            if key in tr:  # ncu dram bytes of one k_pcg_solve launch / its PCG iterations, scaled to this run's average launch
                roofline["traffic"] = tr[key] * k_total / launches_timed
                roofline["traffic_per_pcg_iteration"] = tr[key]
        except Exception:
            pass

    if args.device_only:
        if rank == 0:
            emit({"tuning_only": True, "value": value, "ms_per_step": 1e3 * seconds / max(steps_done, 1), "n_gpus": world,
                  "super_tile_obs": args.super_tile_obs, "schur_mode": args.schur_mode, "workload": args.workload, "structure": info, "parity": parity,
                  "stages_ms_per_step": {k[8:]: 1e3 * v / max(steps_done, 1) for k, v in res.items()
                                         if k.startswith("seconds_") and k != "seconds_total"},
                  "pcg_us_by_phase": [1e6 * v / n_prod for v in res["pcg_phase_seconds"]], "ms_product_phase": prod_ms,
                  "pcg_ms_per_iteration": 1e3 * pcg_seconds / k_total})
        if rank == 0:
            sampler.stop()
        P.close()
        ctx.close()
        if dist is not None:
            dist.destroy_process_group()
        return
    # ---- e2e arm: every step copies its inputs from pinned host memory and reads the result back ----------
    P.set_vertices_raw(h_cams.data_ptr(), h_pts.data_ptr())
    tw, rw = P.lm(iterations=args.warmup, solver=args.solver, schur_mode=args.schur_mode)
    # state after warm-up becomes the host-side state the steps start from
    P.get_vertices_raw(out_cams.data_ptr(), out_pts.data_ptr())
    mu, nu = rw["final_damping"], rw["final_nu"]
    h2d = h_obs.numel() * h_obs.element_size() + out_cams.numel() * out_cams.element_size() + out_pts.numel() * out_pts.element_size()
    d2h = out_cams.numel() * out_cams.element_size() + out_pts.numel() * out_pts.element_size() + 8
    barrier()
    sampler.mark(True)
    t0 = time.perf_counter()
    e2e_steps = 0
    for _ in range(args.steps):
        P.set_observations_raw(h_obs.data_ptr())                       # H2D
        P.set_vertices_raw(out_cams.data_ptr(), out_pts.data_ptr())    # H2D
        tj, rj = P.lm(iterations=1, initial_damping=mu, initial_nu=nu, solver=args.solver, schur_mode=args.schur_mode)  # linearize + one LM iteration
        P.get_vertices_raw(out_cams.data_ptr(), out_pts.data_ptr())    # D2H (+ chi2 in rj)
        mu, nu = rj["final_damping"], rj["final_nu"]
        e2e_steps += 1
    barrier()
    sampler.mark(False)
    serial_seconds = max_over_ranks(time.perf_counter() - t0)

    # the same steps with the observation upload double-buffered on the library's copy stream: the batch of step k+1 is
    # copied (same bytes, inside the timed region) while step k computes; vertices stay in line (step k+1 needs step k's)
    P.set_vertices_raw(h_cams.data_ptr(), h_pts.data_ptr())
    tw, rw = P.lm(iterations=args.warmup, solver=args.solver, schur_mode=args.schur_mode)
    P.get_vertices_raw(out_cams.data_ptr(), out_pts.data_ptr())
    mu, nu = rw["final_damping"], rw["final_nu"]
    P.stage_observations_async(h_obs.data_ptr(), 0)  # batch of step 0
    barrier()
    sampler.mark(True)
    t0 = time.perf_counter()
    for k in range(args.steps):
        P.commit_observations(k % 2)                                     # install this step's batch (copy issued a step ago)
        P.set_vertices_raw(out_cams.data_ptr(), out_pts.data_ptr())      # H2D (first: the step cannot start without them)
        P.stage_observations_async(h_obs.data_ptr(), (k + 1) % 2)        # H2D of the next step's batch, overlaps this step
        tj, rj = P.lm(iterations=1, initial_damping=mu, initial_nu=nu,   # linearize + one LM iteration; the linearisation
                      defer_final_linearize=True, solver=args.solver, schur_mode=args.schur_mode)   # at the accepted point is the next step's first act
        P.get_vertices_raw(out_cams.data_ptr(), out_pts.data_ptr())      # D2H (+ chi2 in rj)
        mu, nu = rj["final_damping"], rj["final_nu"]
    barrier()
    sampler.mark(False)
    e2e_seconds = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else None
    # ONE protocol is the value at every N: the overlapped one (what a streaming caller of this C ABI does); the serial
    # protocol (every copy in line, eager re-linearisation as the reference does) is listed next to it
    overlapped_seconds = e2e_seconds
    e2e = {"value": e2e_steps / overlapped_seconds, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "ms_per_step": 1e3 * overlapped_seconds / max(e2e_steps, 1),
           "protocol": "overlapped",
           "serial_value": e2e_steps / serial_seconds,
           "note": "per step: pinned-host -> device copy of observations + vertices, gb_lm(1 iteration) through the C ABI, "
                   "device -> host copy of the vertices and the cost.  value: the observation batch of step k+1 is "
                   "uploaded on the copy stream (gb_stage_observations_async / gb_commit_observations, double-buffered) "
                   "while step k computes, and an accepted step is not re-linearised before returning (the next step "
                   "linearises the uploaded vertices anyway: gb_lm_options.defer_final_linearize); serial_value: every "
                   "copy in line (gb_set_observations) and the reference's eager re-linearisation"}

    # ---- CPU baseline (rank 0, N=1 only): the oracle port on the host cores, bounded sample -----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.binding import Oracle, default_options
        cores = os.cpu_count() or 1
        O = Oracle(prob, "f64" if tname == "f64" else "f32")
        O.lm_begin(default_options(iterations=args.cpu_baseline_steps + 1, threads=cores))
        O.lm_step()
        t0 = time.perf_counter()
        for _ in range(args.cpu_baseline_steps):
            O.lm_step()
        dtc = time.perf_counter() - t0
        cpu = {"value": args.cpu_baseline_steps / dtc, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"LM iterations 2..{args.cpu_baseline_steps + 1} of the same workload from the same initial state "
                         f"(oracle/oracle_bal.cpp, explicit Schur + PCG, OpenMP, {dtc:.1f} s)"}

    ref_gpu = None
    if rank == 0 and world == 1 and not args.no_reference_gpu:
        ref_gpu = reference_gpu_leg(prob, args.precision, args.solver, args.warmup, steps_done)

    # whole LM iteration against the HBM roofline: SURVEY 8(d)'s algorithmic bytes (per rank) for the PCG iterations and
    # re-linearisations this run actually executed, divided by the measured step time
    k_avg = res["pcg_iterations_total"] / max(steps_done, 1)
    acc_frac = res["accepted"] / max(steps_done, 1)
    step_bytes = lm_iteration_bytes(nc, local.n_pts, info["n_obs"], sT, sS, k_avg, acc_frac)
    step_gbps = step_bytes / (seconds / max(steps_done, 1)) / 1e9
    # the same with the bytes THIS layout has to move (J stored as 24 values per observation instead of
    # the 27 of the E blocks, Jacobian-free back-substitution from the kept point sums): the stricter of the two fractions
    m_, np_, R_, nt_ = info["n_obs"], local.n_pts, info["n_partial_rows"], info["n_tiles"]
    nch_ = info.get("n_camera_chunks", 0)
    lay_lin = m_ * (2 * sT + 4) + 3 * np_ * sT + m_ * (24 * sS + 2 * sT) + 9 * np_ * sT + 18 * R_ * sT
    lay_prep = np_ * 24 * sT + m_ * (24 * sS + 8) + 9 * m_ * sT + 54 * nch_ * sT
    lay_back = (3 * k_avg + 20) * np_ * sT if k_avg <= 16 else m_ * (24 * sS + 8) + 20 * np_ * sT
    lay_cost = m_ * (2 * sT + 8)
    layout_bytes = acc_frac * lay_lin + lay_prep + k_avg * prod_bytes + lay_back + lay_cost
    layout_gbps = layout_bytes / (seconds / max(steps_done, 1)) / 1e9
    lm_roofline = {"bytes_per_step": step_bytes, "pcg_iterations_per_step": k_avg, "accepted_fraction": acc_frac,
                   "achieved": step_gbps, "peak": peak, "unit": "GB/s", "frac": step_gbps / peak,
                   "layout_bytes_per_step": layout_bytes, "layout_achieved": layout_gbps, "layout_frac": layout_gbps / peak,
                   "note": "SURVEY 8(d) algorithmic bytes per LM iteration (implicit Schur), per rank, with the executed PCG "
                           "iterations; at 10 PCG iterations and every step accepted the same formula gives 15.8 GB (Venice FP64)"}
    if args.solver != "pcg-schur":
        # Full-system PCG (PCGSolver, solver/pcg.hpp): the loop is device-resident (no host reads inside a solve), so the CUDA
        # events around the solve time its iterations.  Algorithmic bytes of ONE iteration in this layout: the J^T J product
        # (Jacobians + meta + tile tables, the direction's point part in and the point sums out, camera rows) plus the
        # vector kernels over the 9 Nc + 3 Np unknowns (direction 5 passes, v2 / p.v2 6, x / r update 7, preconditioner 4).
        dimH = 9 * nc + 3 * local.n_pts
        full_bytes = (info["n_obs"] * (24 * sS + 4) + info["n_tiles"] * (2400 - 1024) + local.n_pts * 9 * sT
                      + 2 * info["n_partial_rows"] * 9 * sT + 22 * dimH * sT)
        ach = full_bytes * k_total / pcg_seconds / 1e9
        roofline = {"bound": "hbm", "kernel": "full-system PCG iteration (k_schur_product2<FULL> on the TMA pipeline + the vector "
                                               "kernels of the device-resident loop), CUDA events around each solve",
                    "achieved": ach, "peak": peak, "peak_kind": peak_kind + " (MEASURED_PEAKS.json hbm_gbs, burst copy)",
                    "unit": "GB/s", "frac": ach / peak, "traffic": None, "bytes_per_pcg_iteration": full_bytes,
                    "pcg_iterations_timed": k_total, "ms_per_pcg_iteration": 1e3 * pcg_seconds / k_total,
                    "share_of_step": pcg_seconds / max(res["seconds_total"], 1e-12)}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps_done, "warmup": args.warmup,
            "ms_per_step": 1e3 * seconds / max(steps_done, 1), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64" if tname == "f64" else "f32", "data": "synthetic",
            "config": dict(workload_config(prob, args.precision, world, args.schur_mode), solver=args.solver),
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "lm_roofline": lm_roofline,
            "cpu_baseline": cpu, "reference_gpu": ref_gpu, "parity": parity,
            "structure_seconds": structure_seconds,
            "pcg": {"iterations_per_step": [int(v) for v in traj[:, 3]], "total": int(res["pcg_iterations_total"]),
                    # event-timed: the solve kernel's time / executed PCG iterations (includes its start and its last barrier)
                    "ms_per_iteration": 1e3 * pcg_seconds / k_total,
                    "iteration_gbps": survey_k4 * k_total / pcg_seconds / 1e9,
                    "iteration_frac_of_hbm_peak": survey_k4 * k_total / pcg_seconds / 1e9 / peak,
                    # in-kernel stamps: product phase and the rest of an iteration (row sums, exchange, updates, barrier B)
                    "ms_product_phase": prod_ms, "product_phase_gbps": prod_gbps,
                    "ms_per_iteration_rest": 1e3 * res["update_seconds"] / n_prod,
                    "us_per_iteration_by_phase": dict(zip(["product_cta0", "barrier_A", "sums_exchange", "rows_update", "barrier_B", "of_sums_exchange_until_pushed"],
                                                          [1e6 * v / n_prod for v in res["pcg_phase_seconds"][:6]]))},
            "stages_ms_per_step": {k[8:]: 1e3 * v / max(steps_done, 1) for k, v in res.items()
                                   if k.startswith("seconds_") and k != "seconds_total"},
            "accepted": int(res["accepted"]), "rejected": int(res["rejected"]),
            "chi2": {"start": float(traj[0, 0]) if len(traj) else None, "end": float(res["final_chi2"])},
            "wall_seconds_timed_region": wall, "structure": info,
        }
        emit(line)
    P.close()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    if rank == 0 and parity.get("ok") is False:
        sys.stderr.write(f"bench.py: PARITY MISMATCH against {parity['golden']}: {parity}\n")
        sys.exit(3)


if __name__ == "__main__":
    main()
