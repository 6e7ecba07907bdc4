#!/usr/bin/env python
"""bench.py -- IMEX steps/sec of the PECS per-step hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repository (CUDA path through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference path (oracle port)

A step = one pass of the loop body of the reference's time loop (reference source/SolarCell.cpp:2055-2080): both
carrier RHS assemblies, the four carrier solves, the Poisson RHS and the Poisson solve, on synthetic input: the
reference's default geometry refined to ~1 M DoF per carrier (cfg3: global refinements 7, local 1; 983 040 DoF per
carrier, Poisson 492 800 DoF).  With N > 1 every rank runs its own context at its own applied bias (the I-V sweep
of BASELINE.json config 5): weak scaling, no data-path collective.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "IMEX steps/sec at ~1M DoF/carrier"
UNIT = "steps/s"


def source_sha16(files):
    import hashlib
    h = hashlib.sha256()
    for rel in files:
        with open(os.path.join(ROOT, rel), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def measured_traffic():
    """DRAM bytes per step of the solve kernels from the committed ncu pass (profiles/r02_solve_traffic.json, written by
    scripts/summarize_launches.py) -- or None when that pass was taken from OTHER kernel sources than the ones in the
    tree (the file carries their hash): a stale figure is not reported."""
    path = os.path.join(ROOT, "profiles", "r02_solve_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        if d.get("kernel_source_sha16") == source_sha16(["pecs_b200/csrc/cuda/solve_kernels.cu",
                                                          "pecs_b200/csrc/cuda/solve_kernels.cuh",
                                                          "pecs_b200/csrc/cuda/schur_kernels.cu"]):
            return d.get("dram_bytes_per_step")
    return None


def measured_rhs_traffic():
    """DRAM bytes of one launch of the carrier RHS kernel from the committed ncu capture, or None (same hash rule)"""
    path = os.path.join(ROOT, "profiles", "r02_rhs_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        if d.get("kernel_source_sha16") == source_sha16(["pecs_b200/csrc/cuda/rhs_kernels.cu", "pecs_b200/csrc/rhs_math.hpp"]):
            return d.get("dram_bytes_per_launch")
    return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.lines, self.proc = device, [], None
        self.nvml, self.samples, self.stop_flag, self.thread = None, [], False, None

    # NVML in-process (a query takes well under a millisecond, so a 0.1 s timed region still gets tens of samples);
    # nvidia-smi -lms as the fallback
    def _nvml_loop(self, handle):
        nv = self.nvml
        while not self.stop_flag:
            try:
                self.samples.append((nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM),
                                     nv.nvmlDeviceGetCurrentClocksEventReasons(handle)))
            except Exception:
                break
            time.sleep(0.004)

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            index = self.device
            if visible and all(t.strip().isdigit() for t in visible.split(",")):
                index = int(visible.split(",")[self.device])
            handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml, self.handle = pynvml, handle
            self.thread = threading.Thread(target=self._nvml_loop, args=(handle,), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            nv = self.nvml
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            sm = sorted(c for c, _ in self.samples)
            bits = 0
            for _, r in self.samples:
                bits |= r
            names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
                     "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
                     "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
            try:
                smax = float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM))
            except Exception:
                smax = None
            load = sm[len(sm) // 2:] if sm else []
            return {"sm_mhz": float(load[len(load) // 2]) if load else None, "sm_max_mhz": smax,
                    "reasons": sorted(k for k, v in names.items() if bits & v), "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for line in self.lines:
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                smax = float(p[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # samples under load: the upper half of the clock readings
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi"}


CONIC = False  # --workload cfg4: BASELINE.json configs[3]


def workload_name(g, l):
    if CONIC:
        return (f"conic nanowire (radius one 0.2, radius two 0.6), illumination on, global refinements {g}, local "
                f"refinements {l} (2:1-smoothed interface refinement), fp64, degree 1 (cfg4)")
    return (f"default input_file.prm geometry, global refinements {g}, local refinements {l}, fp64, degree 1" +
            (" (cfg3: 983040 DoF per carrier, Poisson 492800 DoF)" if (g, l) == (7, 1) else ""))


def workload_overrides():
    return {"mesh__radius_one": 0.2, "mesh__radius_two": 0.6} if CONIC else {}


# ------------------------------------------------------------------------------------------------- CPU arm
def cpu_scaling_samples(g_full, l, steps, threads=None):
    """The CPU path (oracle/cpu_arm.py: oracle assembly + SuperLU solves, no product library involved) measured on the
    two refinements below the workload's; returns the samples and the measured exponent of seconds/step in cells."""
    import math
    from oracle import cpu_arm
    samples = [cpu_arm.run(g, l, steps, 1, threads, workload_prm()) for g in (g_full - 2, g_full - 1)]
    a, b = samples
    exponent = math.log(b["seconds_per_step"] / a["seconds_per_step"]) / math.log(b["cells_per_subdomain"] / a["cells_per_subdomain"])
    return samples, exponent


def workload_prm():
    return {"radius one": 0.2, "radius two": 0.6} if CONIC else {}


def run_reference_arm(args):
    """`--impl reference`: the reference's CPU path on the box's host cores ON THE WORKLOAD IT NAMES.  The two smaller
    refinements are always measured (they give the exponent and the cost estimate); the workload itself is run when the
    estimated factorisation fits --cpu-budget-seconds and the free memory, and then `same_config` is true.  Rank 0 alone
    works; under torchrun the value is the one-job figure UNCHANGED (N bias points on one CPU box take N times as long:
    the whole-job steps/s of the box do not grow with N)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0  # only rank 0 runs and prints the CPU arm
    from oracle import cpu_arm
    threads = os.cpu_count()  # never torchrun's OMP_NUM_THREADS=1
    g_full, l = args.global_refinements, args.local_refinements
    t_start = time.perf_counter()
    samples, exponent = cpu_scaling_samples(g_full, l, max(3, min(args.steps, 10)), threads)
    big = samples[-1]
    ratio = cpu_arm.cells_per_subdomain(g_full, l) / big["cells_per_subdomain"]
    growth = big["setup_seconds"]["factor"] / max(samples[0]["setup_seconds"]["factor"], 1e-9)
    est_factor = big["setup_seconds"]["factor"] * growth
    est_step = big["seconds_per_step"] * ratio ** exponent
    nnz_growth = sum(big["factor_nnz"]) / sum(samples[0]["factor_nnz"])
    est_gb = sum(big["factor_nnz"]) * nnz_growth * 12 * 3 / 2 ** 30  # factors + SuperLU work space while factorising
    free_gb = cpu_arm.available_memory_gb()
    steps = args.steps
    fits = (est_factor * 1.5 + est_step * (steps + 1) <= args.cpu_budget_seconds - (time.perf_counter() - t_start)) and \
           (free_gb is None or est_gb <= 0.8 * free_gb)
    if not fits:  # fewer timed steps before giving the configuration up
        steps = max(3, int((args.cpu_budget_seconds - (time.perf_counter() - t_start) - est_factor * 1.5) / max(est_step, 1e-9)) - 1)
        fits = steps >= 3 and est_factor * 1.5 <= args.cpu_budget_seconds and (free_gb is None or est_gb <= 0.8 * free_gb)
        steps = min(max(steps, 3), args.steps)
    estimate = {"factor_seconds": est_factor, "seconds_per_step": est_step, "memory_gb": est_gb, "free_memory_gb": free_gb,
                "budget_seconds": args.cpu_budget_seconds}
    if fits and not args.cpu_no_full:
        full = cpu_arm.run(g_full, l, steps, max(args.warmup, 1), threads, workload_prm())
        sps, same, used = full["steps_per_s"], True, full
        sample = (f"the workload itself: oracle assembly (C++/OpenMP, {threads} threads) + SuperLU (scipy splu, COLAMD; 4 "
                  f"concurrent substitutions + 1) as the UMFPACK stand-in on global refinements {g_full} "
                  f"({full['dofs_per_carrier']} DoF/carrier), {steps} timed steps; factorisation (untimed, like the "
                  f"reference's set_solvers) {full['setup_seconds']['factor']:.0f} s")
    else:
        sps, same, used = 1.0 / est_step, False, big
        steps = big["steps"]
        sample = (f"oracle assembly + SuperLU on global refinements {g_full - 2} and {g_full - 1}, extrapolated to the "
                  f"workload with the MEASURED exponent {exponent:.3f} of seconds/step in cells (the workload itself was "
                  f"estimated at {est_factor:.0f} s of factorisation and {est_gb:.0f} GB: outside the budget)")
    line = {"impl": "reference", "metric": METRIC, "value": sps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 / sps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(g_full, l), "same_config": same,
                       "parallelism": "one CPU box; the figure does not grow with --gpus"},
            "cpu_baseline": {"value": sps, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "same_config": same, "measured_exponent_seconds_per_step_in_cells": exponent,
                             "section_seconds_per_step": used["section_seconds_per_step"],
                             "setup_seconds": used["setup_seconds"], "estimate_for_workload": estimate,
                             "samples": [{k: v[k] for k in ("g", "dofs_per_carrier", "steps_per_s", "setup_seconds")}
                                         for v in samples]},
            "e2e": {"value": sps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------- GPU arm
def run_gpu_arm(args):
    import numpy as np
    import pecs_b200 as pecs
    from pecs_b200 import solarcell as sc, sweep

    rank, local, world, dist = sweep.init_distributed("nccl")
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    device = local
    pinned_cores = None
    if dist is not None:
        import torch
        torch.cuda.set_device(device)
    if pecs.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device; pecs_b200 has no CPU fallback")
    g, l = args.global_refinements, args.local_refinements
    overrides = workload_overrides()
    if world > 1:  # one applied bias per rank; the bias only acts through Dirichlet faces at x == 0 (insulated=false)
        overrides.update({"physical__insulated": False, "physical__applied_bias": sweep.bias_for_rank(rank, world)})
    t_setup = time.perf_counter()
    prob = pecs.SolarCellProblem(pecs.default_input_file(g, l, **overrides), device=device)
    prob.setup_full_system()
    prob.synchronize()
    t_setup = time.perf_counter() - t_setup

    K, W = args.steps, max(args.warmup, 3)
    prob.step(W)
    prob.synchronize()
    sweep.barrier(dist, device)
    sampler = ClockSampler(device)
    sampler.start()
    ms = prob.step_timed(K)[0]  # CUDA events on the context's stream around K graph replays
    clocks = sampler.stop()
    sweep.barrier(dist, device)
    ms_max = sweep.max_over_ranks(ms, dist, device)
    value = world * K / (ms_max * 1e-3)

    # ---- end to end through the host-buffer entry point: H2D of the five states, one step, D2H of the five states
    if dist is not None:  # N > 1: every rank on cores of its GPU's NUMA node (after the multi-threaded setup, before the
        pinned_cores = sweep.pin_to_gpu_numa_node(local, world)  # page-locked buffers are allocated and first touched)
    states = prob.pinned_states()
    for w in range(5):
        states[w][:] = prob.get_solution(w)
    prob.step_host(1, states)  # warm
    sweep.barrier(dist, device)
    t0 = time.perf_counter()
    for _ in range(K):
        prob.step_host(1, states)
    e2e_s = time.perf_counter() - t0
    e2e_s = sweep.max_over_ranks(e2e_s, dist, device)
    h2d_bytes, d2h_bytes = prob.info(sc.INFO_HOST_STEP_H2D_BYTES), prob.info(sc.INFO_HOST_STEP_D2H_BYTES)

    # ---- the benchmarked configuration is a VALIDATED configuration (outside every timed region): the states after the
    # timed steps are finite, and one pass of the hot path from a perturbed state (seed 1234) is checked on THIS mesh:
    # five right-hand sides against the CPU oracle's assembly (1e-12), five solves through |b - A x| / |b| with the
    # host CSR matrices (tests/helpers.py: validate_workload).  The oracle is the checker here, nothing it computes is
    # timed or returned.
    finite_after_timed = bool(all(np.isfinite(prob.get_solution(w)).all() for w in range(5)))
    parity = None
    if rank == 0 and not args.no_validate:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from helpers import validate_workload
        parity = validate_workload(prob)
        parity["tolerance"] = {"rhs_rel": 1e-12, "solve_density_err": 1e-9, "solve_potential_err": 1e-9,
                               "solve_field_err": 1e-9, "solve_current_err": 1e-7}
        parity["finite_after_timed_steps"] = finite_after_timed
        parity["solve_wait_errors"] = prob.info(sc.INFO_SOLVE_WAIT_ERRORS)
        parity["ok"] = bool(parity["rhs_rel"] <= 1e-12 and parity["finite"] and finite_after_timed and
                            parity["solve_density_err"] <= 1e-9 and parity["solve_potential_err"] <= 1e-9 and
                            parity["solve_field_err"] <= 1e-9 and parity["solve_current_err"] <= 1e-7 and
                            parity["solve_wait_errors"] == 0)
        parity["what"] = ("THIS workload. last_step_residual_rel: |b - A x|_inf / |b|_inf of the five systems of the last "
                          "timed step (host CSR mat-vec).  From a perturbed state (seed 1234): rhs_rel = worst block of the "
                          "five assembled right-hand sides vs the CPU oracle; residual_rel / backward_err = residuals of "
                          "the five solves; solve_*_err = error of that ONE solve per block against the solution refined "
                          "with long-double residuals (tests/helpers.py: validate_workload), max-norm relative")
    sweep.barrier(dist, device)

    line = None
    # ---- roofline of the dominant kernels: device time of the solves INSIDE a step = step graph - assembly-only graph
    # (the assembly passes run strictly before / between the solves; both timed with CUDA events on the launching
    # stream over K replays); the sectioned run (one launch at a time, includes launch gaps) only reports the
    # reference's five TimerOutput sections
    rhs_only_ms = prob.step_timed(K, sectioned=3)[0] / K
    solve_ms = sweep.max_over_ranks(ms / K - rhs_only_ms, dist, device)
    sect = prob.step_timed(min(K, 10), sectioned=True) / min(K, 10)
    factor_bytes = prob.info(sc.INFO_FACTOR_BYTES)
    solve_bytes = prob.info(sc.INFO_SOLVE_BYTES_PER_STEP)
    n_solve_launches = prob.info(sc.INFO_LAUNCHES_PER_STEP) - 6
    rhs_ms, rhs_launches = prob.time_kernel(0, 20)
    # the production carrier kernels side by side (0 point-by-point, 1 sum-factorised = default, 2 streaming, 1x shapes)
    rhs_variants = {}
    for v in ("0", "1", "11", "14", "15", "2"):
        os.environ["PECS_B200_RHS_KERNEL"] = v
        rhs_variants[v] = prob.time_kernel(0, 20)[0]
    del os.environ["PECS_B200_RHS_KERNEL"]
    prhs_ms = prob.time_kernel(1, 20)[0]
    rhs_bytes = 368 * (prob.n_cells(0) + prob.n_cells(1))
    peak, peak_src = measured_peaks()
    if rank == 0:
        solve_gbs = solve_bytes / (solve_ms * 1e-3) / 1e9
        rhs_gbs = rhs_bytes / (rhs_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(g, l), "parallelism": "1 context per GPU, one applied bias per rank"
                       if world > 1 else "single context",
                       "l2": "inputs larger than L2: every step streams the factor tables "
                             f"({factor_bytes / 1e9:.1f} GB) once; the isolated RHS timing flushes L2 (256 MB memset + 256 MB read sweep, so evictions are clean)",
                       "setup_seconds": t_setup,
                       "host_cores_of_rank_0": (f"{pinned_cores[0]}-{pinned_cores[-1]}" if pinned_cores else "not pinned")},
            "e2e": {"value": world * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes,
                    "what": "pecs_step_host on pinned host states: upload of what a step reads (4 density blocks + "
                            "Poisson vector), one step, download of all five state vectors (overlapped per species)"},
            "gpu_launches": prob.info(sc.INFO_LAUNCHES_PER_STEP) * K,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "forward/backward level kernels (+ residual/recover ELL kernels) of "
                                                   "the five multifrontal solves of one step",
                         "achieved": solve_gbs, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": solve_gbs / peak, "traffic": measured_traffic(),
                         "algorithmic_bytes_per_step": solve_bytes, "launches_per_step": n_solve_launches,
                         "ms_per_step": solve_ms,
                         "note": "achieved = algorithmic bytes of one step's solves (8 B per front-operator entry, 12 B per "
                                 "ELL entry, each read once) / device time of the solves inside the step graph (step graph "
                                 "minus assembly-only graph, CUDA events); traffic = DRAM bytes read+written by the same "
                                 "kernels in the committed ncu pass (profiles/)"},
            "rhs_roofline": {"bound": "hbm", "kernel": "carrier_rhs_direct_kernel: cell + boundary terms of all four "
                                                       "carriers, both subdomains, one launch, L2 flushed before it",
                             "achieved": rhs_gbs, "peak": peak, "unit": "GB/s", "frac": rhs_gbs / peak,
                             "algorithmic_bytes_per_launch_group": rhs_bytes, "ms": rhs_ms, "launches": rhs_launches,
                             "traffic": measured_rhs_traffic(),
                             "variants_ms": rhs_variants,
                             "poisson_rhs_ms": prhs_ms,
                             "poisson_rhs_gbs": 140 * (prob.n_cells(0) + prob.n_cells(1)) / (prhs_ms * 1e-3) / 1e9},
            "section_ms_per_step": dict(zip(["Assemble semiconductor rhs", "Assemble electrolyte rhs",
                                             "Solve LDG Systems", "Assemble Poisson rhs", "Solve Poisson system"],
                                            [float(x) for x in sect[1:]])),
            "section_note": "sections are timed one call at a time (no overlap between sections, launch gaps included); "
                            "their sum exceeds ms_per_step, which is the captured graph",
            "parity": parity,
        }
    prob.close()
    # ---- N = 2 / 4: the COMMUNICATING layouts of the same workload under the same clock (strong scaling of one step:
    # 2 ranks by subdomain, 4 by species; density exchange fused into the solves over peer memory).  The sweep stays `value`.
    if world in (2, 4) and not args.no_sharded:
        sh, _ = measure_sharded(args, rank, device, world, dist, args.exchange)
        if rank == 0:
            one_gpu = value / world  # steps/s of one context on one GPU, measured a moment ago in this very run
            sh["speedup_vs_one_gpu"] = sh["steps_per_s"] / one_gpu
            sh["strong_scaling_efficiency"] = sh["steps_per_s"] / one_gpu / world
            sh["one_gpu_steps_per_s"] = one_gpu
            line["sharded"] = sh
    # ---- the reference's DEFAULT input end to end (cfg1 = input_file.prm verbatim: g=4, l=1, 1000 steps, 100 time stamps,
    # reference main.cpp -> run_full_system): setup + time loop + the 101 output stamps + restart files, wall clock
    if rank == 0 and world == 1 and not args.no_cfg1:
        # (side measurements: a failure here -- a full disk, say -- must not cost the line its headline numbers)
        try:
            line["cfg1_default_input"] = measure_default_input(device, args)
        except Exception as e:  # noqa: BLE001
            line["cfg1_default_input"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        # ---- and the same main() on the benchmarked workload: setup + 1000 steps + the time stamps as .vtu files
        try:
            line["workload_run_full_system"] = measure_workload_main(device, args)
        except Exception as e:  # noqa: BLE001
            line["workload_run_full_system"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    if rank == 0:
        # CPU baseline on this box's host cores (rank 0, N = 1 only): bounded sample of the same workload (the two
        # refinements below it, MEASURED exponent); `--impl reference` runs the workload itself
        if world == 1 and not args.no_cpu_baseline:
            samples, exponent = cpu_scaling_samples(g, l, args.cpu_steps, os.cpu_count())
            big = samples[-1]
            ratio = prob_cells(g, l) / big["cells_per_subdomain"]
            line["cpu_baseline"] = {
                "value": 1.0 / (big["seconds_per_step"] * ratio ** exponent), "unit": UNIT, "cores": big["threads"],
                "kind": "port",
                "sample": f"oracle assembly (C++/OpenMP) + SuperLU substitutions (UMFPACK stand-in, 4 concurrent + 1) on "
                          f"global refinements {g - 2} and {g - 1} ({args.cpu_steps} steps each), extrapolated to the "
                          f"workload with the measured exponent {exponent:.3f} of seconds/step in cells; "
                          f"`--impl reference` times the workload itself",
                "measured": [{k: v[k] for k in ("g", "dofs_per_carrier", "steps_per_s")} for v in samples]}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def measure_default_input(device, args):
    """BASELINE.json configs[0] on the GPU path: `run_full_system` of input_file.prm as the reference's main() runs it."""
    import shutil
    import tempfile
    import pecs_b200 as pecs
    out = {"workload": "input_file.prm verbatim: global refinements 4, local 1 (15360 DoF per carrier), 1000 IMEX steps, "
                       "100 time stamps (303 .vtu files + 4 restart files)"}
    tmp = tempfile.mkdtemp(prefix="pecs_cfg1_")
    try:
        t0 = time.perf_counter()
        prob = pecs.SolarCellProblem(pecs.default_input_file(4, 1), device=device)
        prob.set_output(tmp)
        prob.run_full_system()
        out["run_full_system_wall_seconds"] = time.perf_counter() - t0
        out["files_written"] = len(os.listdir(tmp))
        ms = prob.step_timed(1000)[0]
        out["steps_per_s_time_loop_only"] = 1000 / (ms * 1e-3)
        out["ms_per_step"] = ms / 1000
        out["gpu_launches_per_step"] = prob.info(0)
        out["finite"] = bool(all(__import__("numpy").isfinite(prob.get_solution(w)).all() for w in range(5)))
        prob.close()
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    if not args.no_cpu_baseline:
        from oracle import cpu_arm
        t0 = time.perf_counter()
        cpu = cpu_arm.run(4, 1, 200, 1, os.cpu_count())
        out["cpu_steps_per_s"] = cpu["steps_per_s"]
        out["cpu_note"] = (f"oracle assembly + SuperLU on {cpu['threads']} host threads, 200 timed steps of the same input "
                           f"(time loop only, no output), setup {sum(cpu['setup_seconds'].values()):.1f} s")
    return out


def measure_workload_main(device, args):
    """`run_full_system` of the benchmarked workload, wall clock: one-time setup (a second context of this process: no
    CUDA start-up in it), 1000 IMEX steps, the reference's 101 time stamps written as .vtu (135 MB each at cfg3) and the
    restart files.  The files go to tmpfs when there is room for them (14 GB), else 10 stamps go to the temp directory."""
    import shutil
    import tempfile
    import pecs_b200 as pecs
    g, l = args.global_refinements, args.local_refinements
    where, stamps = tempfile.gettempdir(), 10
    try:
        if shutil.disk_usage("/dev/shm").free > 40e9:
            where, stamps = "/dev/shm", 100
    except OSError:
        pass
    tmp = tempfile.mkdtemp(prefix="pecs_main_", dir=where)
    out = {"workload": workload_name(g, l), "time_stamps": stamps, "directory": where}
    try:
        prob = pecs.SolarCellProblem(pecs.default_input_file(g, l, computational__time_stamps=stamps, **workload_overrides()),
                                     device=device)
        prob.set_output(tmp)
        t0 = time.perf_counter()
        prob.run_full_system()
        out["run_full_system_wall_seconds"] = time.perf_counter() - t0
        files = os.listdir(tmp)
        out["files_written"] = len(files)
        out["gigabytes_written"] = sum(os.path.getsize(os.path.join(tmp, f)) for f in files) / 1e9
        prob.close()
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return out


# ------------------------------------------------------------------------------------------------- sharded step
def measure_sharded(args, rank, device, world, dist, exchange):
    """ONE step spread over 2 (subdomains) or 4 (species) GPUs (pecs_b200/shard.py): strong scaling of the same workload.
    Timed with CUDA events on the context's stream, max over ranks.  Returns (dict for rank 0, clocks)."""
    import torch
    import pecs_b200 as pecs
    from pecs_b200 import shard, solarcell as sc, sweep
    g, l = args.global_refinements, args.local_refinements
    t_setup = time.perf_counter()
    prob = pecs.SolarCellProblem(pecs.default_input_file(g, l, **workload_overrides()), device=device)
    prob.set_owned_species(shard.owned_mask(rank, world))
    prob.setup_full_system()
    prob.synchronize()
    t_setup = time.perf_counter() - t_setup
    engine = shard.GpuEngine(prob, device)
    if exchange == "p2p":
        engine.connect_p2p(dist, rank, world)
    stepper = shard.ShardedStepper(engine, dist, rank, world)
    K, W = args.steps, max(args.warmup, 3)
    stepper.step(W)
    prob.synchronize()
    sweep.barrier(dist, device)
    sampler = ClockSampler(device)
    sampler.start()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(engine.stream):
        a.record()
    stepper.step(K)
    with torch.cuda.stream(engine.stream):
        b.record()
    b.synchronize()
    ms = a.elapsed_time(b)
    clocks = sampler.stop()
    sweep.barrier(dist, device)
    ms_max = sweep.max_over_ranks(ms, dist, device)
    exchanged = sum(prob.density_block(s)[1] for s in range(4)) * 8
    solve_bytes = sweep.sum_over_ranks(prob.info(sc.INFO_SOLVE_BYTES_PER_STEP), dist, device)
    launches = sweep.sum_over_ranks(prob.info(sc.INFO_LAUNCHES_PER_STEP), dist, device)
    wait_errors = sweep.sum_over_ranks(prob.info(sc.INFO_SOLVE_WAIT_ERRORS), dist, device)
    finite = all(bool(torch.isfinite(engine.density(s)).all().item()) for s in range(4))
    finite = sweep.sum_over_ranks(0 if finite else 1, dist, device) == 0
    prob.close()
    out = {"layout": shard.mode_name(world), "steps_per_s": K / (ms_max * 1e-3), "ms_per_step": ms_max / K, "steps": K,
           "exchange": (f"NCCL broadcast of the four density blocks per step ({exchanged} B)" if exchange == "nccl" else
                        f"fused into the backward sweeps: peer-memory stores over NVLink ({exchanged} B per step and "
                        "peer), flag kernels, no collective") + "; Poisson solved on every rank",
           "bytes_exchanged_per_step": int(exchanged) * (world - 1), "solve_bytes_per_step_all_ranks": int(solve_bytes),
           "gpu_launches_per_step_all_ranks": int(launches), "setup_seconds": t_setup, "finite": bool(finite),
           "solve_wait_errors": int(wait_errors)}
    return out, clocks


def run_sharded_arm(args):
    """`--parallelism subdomain|species`: the sharded step as the headline value of the line (strong scaling)"""
    import torch
    from pecs_b200 import sweep

    rank, local, world, dist = sweep.init_distributed("nccl")
    want = {"subdomain": 2, "species": 4}[args.parallelism]
    if world != args.gpus or world != want:
        raise SystemExit(f"--parallelism {args.parallelism} needs exactly {want} ranks (torchrun --nproc-per-node {want})")
    torch.cuda.set_device(local)
    g, l = args.global_refinements, args.local_refinements
    sh, clocks = measure_sharded(args, rank, local, world, dist, args.exchange)
    if rank == 0:
        peak, peak_src = measured_peaks()
        gbs = sh["solve_bytes_per_step_all_ranks"] / (sh["ms_per_step"] * 1e-3) / 1e9
        line = {"metric": METRIC, "value": sh["steps_per_s"], "unit": UNIT, "n_gpus": world, "steps": sh["steps"],
                "warmup": max(args.warmup, 3), "ms_per_step": sh["ms_per_step"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(g, l), "parallelism": sh["layout"], "exchange": sh["exchange"],
                           "l2": "inputs larger than L2: every step streams the factor tables once",
                           "setup_seconds": sh["setup_seconds"]},
                "e2e": None, "gpu_launches": sh["gpu_launches_per_step_all_ranks"] * sh["steps"], "clocks": clocks,
                "roofline": {"bound": "hbm", "kernel": "level kernels of the solves, all ranks", "achieved": gbs,
                             "peak": peak * world, "peak_source": peak_src, "unit": "GB/s", "frac": gbs / (peak * world),
                             "traffic": None, "note": "whole step (incl. exchange and RHS) as denominator; the Poisson "
                                                      "tables are streamed by every rank"},
                "sharded": sh, "cpu_baseline": None}
        print(json.dumps(line))
    dist.destroy_process_group()
    return 0


def prob_cells(g, l):
    return 4 ** g + (4 ** (g + l) if l > 0 else 0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["pecs_b200", "reference"], default="pecs_b200")
    ap.add_argument("--global-refinements", type=int, default=7)
    ap.add_argument("--local-refinements", type=int, default=1)
    ap.add_argument("--cpu-steps", type=int, default=5, help="timed steps of each bounded CPU sample of the GPU arm")
    ap.add_argument("--cpu-budget-seconds", type=float, default=1500.0,
                    help="--impl reference: wall-clock budget; the workload itself is run when its estimated cost fits")
    ap.add_argument("--cpu-no-full", action="store_true", help="--impl reference: samples + extrapolation only")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cfg1", action="store_true", help="skip the end-to-end run of the reference's default input")
    ap.add_argument("--no-sharded", action="store_true", help="N = 2 / 4: skip the sharded-step measurement")
    ap.add_argument("--no-validate", action="store_true", help="skip the parity check of the benchmarked workload")
    ap.add_argument("--exchange", choices=["nccl", "p2p"], default="p2p",
                    help="sharded step: 'p2p' = density exchange fused into the solves over peer memory (default), "
                         "'nccl' = broadcasts between the two halves of the step")
    ap.add_argument("--parallelism", choices=["sweep", "subdomain", "species"], default="sweep",
                    help="N > 1: 'sweep' = one applied bias per rank (weak scaling, default); 'subdomain' (2 ranks) / "
                         "'species' (4 ranks) = one step sharded over the ranks with an NCCL density exchange (strong)")
    ap.add_argument("--workload", choices=["cfg3", "cfg4"], default="cfg3",
                    help="cfg3 = the headline configuration (default); cfg4 = conic wire, two levels of interface "
                         "refinement, global refinements 6 unless given (BASELINE.json configs[3]; a parity-test "
                         "configuration, benchable for the sharded layouts)")
    args = ap.parse_args()
    if args.workload == "cfg4":
        global CONIC
        CONIC = True
        if "--global-refinements" not in sys.argv:
            args.global_refinements = 6
        if "--local-refinements" not in sys.argv:
            args.local_refinements = 2
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.parallelism != "sweep":
        return run_sharded_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
