#!/usr/bin/env python
"""bench.py — throughput of the path-tracing hot path on B200 (Mpaths/s), next to the CPU oracle.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line.
A "step" is one full frame of the workload: BASELINE.json's headline config C5 = book-2 `final` scene, 800x800,
10 000 spp, depth 50 (6.4 G paths).  For N > 1 (torchrun, one rank per GPU) the samples of every pixel are sharded
across ranks (strong scaling: the frame is fixed), one NCCL all-reduce(sum) combines the accumulators, rank 0 resolves.

  value   = whole-job Mpaths/s with the scene resident in HBM: (W*H*spp) / step time; device-timed (CUDA events on
            the launch stream), max over ranks.
  e2e     = the same metric through the public C-ABI calls with HOST buffers: every step uploads the scene tables
            (H2D) and reads the resolved RGBA-f32 frame back (D2H) inside the timed region.
  roofline= FP32 issue: algorithmic flops/path (oracle operation counts x reference per-operation flop costs,
            profiles/work_model.json) x paths/s vs the FFMA peak measured on this GPU by hrt_measure_peaks.
  cpu_baseline = the C++ oracle (a port: the Rust reference cannot be built here) on all host cores, bounded sample.

`--impl reference` times the oracle port alone (rank 0 only) and prints the same line with "impl": "reference".
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

import __graft_entry__ as graft  # noqa: E402

METRIC = "Mpaths/s, book-2 final scene 800x800"  # BASELINE.json's metric (config C5)
UNIT = "Mpaths/s"


def metric_name(config, scene_name, width, height):
    """BASELINE.json's metric for its headline config; the other configs are labelled with their own scene and size."""
    return METRIC if config == "C5" else f"Mpaths/s, {scene_name} scene {width}x{height}"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C5", help="BASELINE config (C1, C2a, C2b, C3, C4, C5)")
    ap.add_argument("--samples", type=int, default=0, help="override spp of the config")
    ap.add_argument("--seed", type=int, default=1, help="scene-instance seed")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--bvh", default="trees", choices=["trees", "reference"],
                    help="flattened form that renders (include/hrt.h hrt_scene_set_bvh_builder): sound BVHs as OP_BVH trees "
                         "(default), or every BvhNode as the reference built it")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="frames of the e2e leg (default max(2, steps // 8))")
    ap.add_argument("--single-process", action="store_true",
                    help="N GPUs from ONE process through hrt_render_multi (one host thread per device, peers' accumulators summed "
                         "over NVLink peer memory inside the resolve kernel) instead of torchrun + NCCL; prints the same line")
    return ap.parse_args()


# ---- clocks ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return

        def pump():
            for line in self.proc.stdout:
                self.rows.append(line.strip())
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---- work model -------------------------------------------------------------------------------------------------
# Reference-arithmetic flop costs per operation: AABB test 27 (15 arith + 12 compares), sphere test
# ~30 (24 miss / 65 hit), rect test ~12 (4 early-out / 12 bounds-fail / 30 hit), medium query 25 (+ its two boundary
# queries, counted as sphere/rect tests), perlin noise 210, scatter 40, per-ray camera/shading bookkeeping 20.
FLOPS = {"aabb_tests": 27.0, "sphere_tests": 30.0, "rect_tests": 12.0, "medium_queries": 25.0, "noise_evals": 210.0,
         "scatters": 40.0, "rays": 20.0}
BYTES = {"aabb_tests": 32.0, "sphere_tests": 32.0, "rect_tests": 32.0 / 6.0 + 32.0 * 0.0, "medium_queries": 32.0,
         "noise_evals": 8 * 16.0 + 6.0, "scatters": 32.0, "rays": 0.0}


def work_model(config: str, counters=None):
    """flops/path and table bytes/path under the minimum (tight-box, ordered) traversal: from profiles/work_model.json
    (generated by tools/work_model.py from oracle counters), else from live oracle counters."""
    path = os.path.join(ROOT, "profiles", "work_model.json")
    per_path = None
    source = None
    if os.path.exists(path):
        with open(path) as f:
            wm = json.load(f)
        if config in wm:
            per_path = wm[config]["tight_per_path"]
            source = "profiles/work_model.json (oracle counters, tight boxes)"
    if per_path is None and counters is not None:
        per_path = {k: getattr(counters, k) / max(1, counters.paths) for k in FLOPS}
        source = "live oracle counters (reference-loose boxes)"
    if per_path is None:
        return None
    flops = sum(FLOPS[k] * per_path[k] for k in FLOPS)
    nbytes = sum(BYTES[k] * per_path[k] for k in BYTES)
    return {"flops_per_path": flops, "bytes_per_path": nbytes, "per_path": per_path, "source": source}


# ---- CPU arm ------------------------------------------------------------------------------------------------------
def cpu_sample(spec, width, height, depth, budget_s, threads=0):
    """Time the oracle on a bounded sample of the workload: full frame, reduced spp (throughput is spp-independent)."""
    orc = graft.load_oracle()
    pkg = graft.load_package()
    ob = orc.OracleBackend()
    pkg.scene.emit(spec.world, ob)
    cores = threads or (os.cpu_count() or 1)
    _, _, c1 = ob.render(spec.camera, width, height, 1, depth, spec.background, seed=1, threads=cores)
    rate = c1.paths / max(c1.seconds, 1e-6)
    spp = int(max(1, min(64, (budget_s * rate) // (width * height))))
    if spp > 1:
        _, _, c = ob.render(spec.camera, width, height, spp, depth, spec.background, seed=2, threads=cores)
    else:
        c = c1
    return {"value": c.paths / c.seconds / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{width}x{height} full frame at {spp} spp, depth {depth} ({c.paths} paths, {c.seconds:.1f} s)",
            "mrays_per_s": c.rays / c.seconds / 1e6, "rays_per_path": c.rays / max(1, c.paths)}, c


def run_reference(args, scene_name, width, height, samples, depth):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg = graft.load_package()
    spec = pkg.make_scene(scene_name, args.seed)
    per_step = max(4.0, min(20.0, 90.0 / max(1, args.steps + args.warmup)))
    times, paths = [], 0
    last = None
    for i in range(args.warmup + args.steps):
        res, c = cpu_sample(spec, width, height, depth, per_step)
        if i >= args.warmup:
            times.append(c.seconds)
            paths += c.paths
            last = res
    value = paths / sum(times) / 1e6
    line = {"impl": "reference", "metric": metric_name(args.config, scene_name, width, height), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{scene_name} {width}x{height} depth {depth} (bounded sample of the {samples}-spp frame)",
                       "scene_seed": args.seed},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": "port", "sample": last["sample"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---- single-process multi-GPU arm (hrt_render_multi) ------------------------------------------------------------------
def run_single_process(args, scene_name, width, height, samples, depth):
    """The shape the reference needs (it is one process): devices 0..N-1 driven by hrt_render_multi.  Every step is the
    whole public call with HOST output, so `value` and `e2e` are the same measurement (host wall clock around blocking
    calls; each device's render is CUDA-event-timed inside and reported as kernel_ms = the slowest device)."""
    pkg = graft.load_package()
    n = args.gpus
    if pkg.native.device_count() < n:
        raise SystemExit(f"bench.py: {n} CUDA devices needed, {pkg.native.device_count()} visible")
    spec = pkg.make_scene(scene_name, args.seed)
    r = pkg.renderer.Renderer(spec, device=0)
    devs = list(range(n))
    out = np.empty((height, width, 4), dtype=np.float32)
    total_paths = width * height * samples
    for i in range(args.warmup):
        r.backend.render_multi(devs, spec.camera, width, height, samples, depth, spec.background, seed=1000 + i, out=out)
    sampler = ClockSampler(0)
    sampler.start()
    kernel_ms, launches = [], 0
    t0 = time.perf_counter()
    for i in range(args.steps):
        _, st = r.backend.render_multi(devs, spec.camera, width, height, samples, depth, spec.background, seed=i, out=out)
        kernel_ms.append(st.kernel_ms)
        launches += st.launches
    ms_per_step = (time.perf_counter() - t0) * 1e3 / args.steps
    clocks = sampler.stop()
    value = total_paths / (ms_per_step * 1e-3) / 1e6
    chk = 16 * n
    a, _ = r.backend.render_multi(devs, spec.camera, width, height, chk, depth, spec.background, seed=4242)
    b, _ = r.render(width, height, chk, depth, seed=4242)
    fin = np.isfinite(a).all(axis=-1) & np.isfinite(b).all(axis=-1)
    parity = bool(np.allclose(a[fin], b[fin], rtol=3e-4, atol=3e-4))
    line = {"metric": metric_name(args.config, scene_name, width, height), "value": value, "unit": UNIT, "n_gpus": n, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.config}: {scene_name} {width}x{height}, {samples} spp, depth {depth}", "scene_seed": args.seed,
                       "bvh": args.bvh, "parallelism": f"single process, hrt_render_multi over {n} devices (peer-memory reduce + resolve)"},
            "kernel_ms": float(np.mean(kernel_ms)), "gpu_launches": int(launches), "clocks": clocks,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": width * height * 16, "ms_per_step": ms_per_step,
                    "api": "hrt_render_multi (host RGBA-f32 out); scene tables resident"},
            "parity_vs_n1": parity}
    print(json.dumps(line), flush=True)


# ---- B200 arm -----------------------------------------------------------------------------------------------------
def run_b200(args, scene_name, width, height, samples, depth):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    pkg = graft.load_package()
    if pkg.native.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback")
    spec = pkg.make_scene(scene_name, args.seed)
    dr = pkg.renderer.DistributedRenderer(spec, width, height, device=local_rank, rank=rank, world_size=world)
    total_paths = width * height * samples

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize; CUDA events on the launch stream; max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    peaks = pkg.native.measure_peaks(local_rank) if rank == 0 else None

    # warm-up (also first-launch module load, NCCL channel setup)
    for i in range(args.warmup):
        dr.step(samples, depth, seed=1000 + i)
    barrier()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    stats = []

    def resident_step(i):
        st = dr.step(samples, depth, seed=i, want_stats=(rank == 0))
        if st is not None and rank == 0:
            stats.append((st.kernel_ms, st.rays, st.paths, st.grid, st.block, st.launches))
    ms_total = timed(resident_step, args.steps)
    clocks = sampler.stop() if sampler else None
    ms_per_step = ms_total / args.steps
    value = total_paths / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: public API, host buffers, scene upload + frame read-back inside the timed region ----
    e2e = None
    if not args.no_e2e:
        # The e2e leg repeats the same full frames through the host-buffer API; it is timed over fewer of them
        # (max(2, steps // 8)) so that `--steps 20 --warmup 5` fits the driver's per-N wall-clock limit.
        e2e_steps = args.e2e_steps if args.e2e_steps > 0 else max(2, args.steps // 8)
        h2d = dr.r.backend.device_bytes() + 128
        d2h = width * height * 16
        if world == 1:
            out = np.empty((height, width, 4), dtype=np.float32)

            def e2e_step(i):
                dr.r.backend.refresh(local_rank)  # H2D: re-copy the scene tables (op stream, materials, textures, texels)
                dr.r.render(width, height, samples, depth, seed=2000 + i, out=out)  # blocking; D2H of the frame
            t0 = time.perf_counter()
            for i in range(e2e_steps):
                e2e_step(i)
            e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        else:
            def e2e_step(i):
                dr.r.backend.refresh(local_rank)
                dr.step(samples, depth, seed=2000 + i, to_host=True)
                torch.cuda.current_stream().synchronize()
            e2e_ms = timed(e2e_step, e2e_steps) / e2e_steps
        e2e = {"value": total_paths / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms, "steps": e2e_steps,
               "api": "hrt_scene_refresh (H2D tables) + hrt_render (host RGBA-f32 out)" if world == 1 else
                      "hrt_scene_refresh + hrt_render_accum_device + NCCL all_reduce + hrt_resolve_device + D2H"}

    # ---- outside the timed region: the N-rank frame against the 1-rank frame of the same sample set ----
    # (spp sharding + one all-reduce must give the 1-GPU image up to f32 summation order)
    chk_spp = 16 * world
    dr.step(chk_spp, depth, seed=4242, to_host=True)
    torch.cuda.current_stream().synchronize()
    parity = None
    if rank == 0:
        sharded = dr.host.numpy().copy()
        solo, _ = dr.r.render(width, height, chk_spp, depth, seed=4242)
        fin = np.isfinite(sharded).all(axis=-1) & np.isfinite(solo).all(axis=-1)
        diff = np.abs(sharded[fin] - solo[fin])
        parity = {"ok": bool(np.allclose(sharded[fin], solo[fin], rtol=3e-4, atol=3e-4)), "max_abs_diff": float(diff.max()) if diff.size else 0.0,
                  "spp": chk_spp, "what": f"{world}-rank sharded frame vs rank 0 alone, same seed, resolved RGBA"}
    barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    kernel_ms = float(np.mean([s[0] for s in stats])) if stats else None
    rays_per_path = float(np.mean([s[1] / max(1, s[2]) for s in stats])) if stats else None
    my_paths = stats[0][2] if stats else total_paths

    cpu = None
    counters = None
    if not args.no_cpu_baseline:
        cpu, counters = cpu_sample(spec, width, height, depth, args.cpu_seconds)

    wm = work_model(args.config, counters)
    roofline = None
    if wm and kernel_ms:
        achieved = wm["flops_per_path"] * my_paths / (kernel_ms * 1e-3) / 1e12
        l2_ach = wm["bytes_per_path"] * my_paths / (kernel_ms * 1e-3) / 1e9
        hbm_peak = None
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                hbm_peak = json.load(f).get("hbm_gbs")
        except OSError:
            pass
        traffic = None
        issue = None
        try:  # DRAM bytes per launch of the dominant kernel, from a committed ncu metric pass of this very workload
            with open(os.path.join(ROOT, "profiles", "r02b_traffic.json")) as f:
                t = json.load(f).get(args.config)
            if t and world == 1:
                # DRAM bytes per path (ncu, summed over every launch of a short frame) x the paths of one step
                traffic = t["dram_bytes_per_path"] * my_paths
                if "issue_active_pct" in t:  # measured issue-slot utilisation of the same kernel (not algorithmic work)
                    issue = {"issue_active_pct": t["issue_active_pct"], "source": t["issue_source"]}
        except OSError:
            pass
        roofline = {"bound": "fp32_issue", "achieved": achieved, "peak": peaks.fp32_tflops, "unit": "TFLOP/s",
                    "frac": achieved / peaks.fp32_tflops, "traffic": traffic,
                    "peak_source": "measured live by hrt_measure_peaks (FFMA chains, CUDA events); MEASURED_PEAKS.json has "
                                   "no FP32 figure",
                    # hrt_api.cu render_into: the wavefront render for big jobs on scenes with OP_BVH trees (kernel_ms is then the
                    # whole pipeline: wave_logic + wave_noise + wave_tree per tree + wave_trace per iteration; shares in
                    # profiles/r02b_wave_window_*.txt), else the persistent uniform-walk kernel
                    "kernel": "wavefront render (wave_logic / wave_tree / wave_trace kernels)" if stats and stats[0][5] > 2 else "render_interp_kernel<true> (uniform walk)",
                    "kernel_ms": kernel_ms, "flops_per_path": wm["flops_per_path"],
                    "work_model": wm["source"], "issue": issue,
                    "l2": {"achieved_gbs": l2_ach, "peak_gbs": peaks.l2_read_gbs, "frac": l2_ach / peaks.l2_read_gbs,
                           "bytes_per_path": wm["bytes_per_path"]},
                    "hbm": {"achieved_gbs": (width * height * 16 * 2) / (kernel_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                            "note": "accumulator traffic only; the working set is L1/L2-resident"}}

    line = {"metric": metric_name(args.config, scene_name, width, height), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"{args.config}: {scene_name} {width}x{height}, {samples} spp, depth {depth}",
                       "scene_seed": args.seed, "bvh": args.bvh, "parallelism": f"spp-sharded x{world}" if world > 1 else "single GPU",
                       "l2_policy": "working set (<3 MB scene tables + 10 MB accumulator) is cache-resident by design; the "
                                    "kernel is compute/latency-bound, not HBM-bound, so no L2 flush applies",
                       "grid": stats[0][3] if stats else None, "block": stats[0][4] if stats else None},
            "mrays_per_s": (value * rays_per_path) if rays_per_path else None, "rays_per_path": rays_per_path,
            "kernel_ms": kernel_ms, "gpu_launches": int(sum(s[5] for s in stats)) + args.steps, "clocks": clocks, "e2e": e2e,
            "roofline": roofline, "parity_vs_n1": parity["ok"] if parity else None, "parity": parity, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.bvh != "trees":
        os.environ["HRT_BVH_BUILDER"] = args.bvh  # read by hrt_scene_create (diagnostic default of the library)
    pkg = graft.load_package()
    scene_name, width, height, samples, depth = pkg.CONFIGS[args.config]
    if args.samples > 0:
        samples = args.samples
    if args.impl == "reference":
        run_reference(args, scene_name, width, height, samples, depth)
    elif args.single_process:
        run_single_process(args, scene_name, width, height, samples, depth)
    else:
        run_b200(args, scene_name, width, height, samples, depth)


if __name__ == "__main__":
    main()
