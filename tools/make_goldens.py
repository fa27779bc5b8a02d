"""Golden frames for image parity: the CPU oracle's own render of each BASELINE config at 4096 spp, depth 50.

The reference cannot be built or run here (no Rust toolchain; it needs a GLFW window and writes no image), so "the
reference's own render" is the oracle's.  The reference-faithful (loose-box) traversal runs at ~0.09 Mpaths/s on
8 cores for `final`, so goldens are rendered at HALF the config's resolution (same scene instance, same camera,
same spp and depth; a pixel just covers 4x the footprint) — full-res C5 alone would take > 8 h.

Each golden is rendered as two independent 2048-spp halves (seeds A, B): their sum is the 4096-spp golden, their
difference measures the oracle-vs-oracle noise floor.  Stored per config in tests/golden/<cfg>.npz:
  lin    float32 [h,w,3]  LINEAR per-pixel mean  sum/4096  (the gamma-resolved golden is sqrt(lin), application.rs:451-456)
  sigma  float32 [h,w,3]  standard error of the linear per-pixel mean  sqrt(var/4096)
  meta   json: sizes, spp, seeds, MAE(A,B) of the two resolved halves, oracle counters, seconds
Files written by the first version of this script hold `img` = sqrt(lin) and `sigma` as float16 instead (the low-light
configs C3 / C4 / C5, hours of CPU each, are kept in that form): fine for MAE / PSNR, but float16 in gamma space is a
5.6e-4 relative error on the constant (0.7, 0.8, 1.0) background of C1 / C2a / C2b — far above their 4096-spp standard
error, which showed up as pooled z-scores of +2 .. +22 against those goldens; they were regenerated in this form.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
orc = graft.load_oracle()
# GOLDEN_SPP / GOLDEN_DIV / GOLDEN_SUFFIX: "deep" goldens (e.g. 65536 spp at quarter resolution) that show convergence
# beyond 40 dB on the low-light scenes, where 4096 spp of the reference is itself ~35 dB from converged.
SPP = int(os.environ.get("GOLDEN_SPP", "4096"))
DIV = int(os.environ.get("GOLDEN_DIV", "2"))
SUFFIX = os.environ.get("GOLDEN_SUFFIX", "")
threads = int(os.environ.get("GOLDEN_THREADS", "7"))
# the oracle hands out tiles dynamically: small frames need small tiles or most threads idle (80 is the reference's default)
TILE = int(os.environ.get("GOLDEN_TILE", "80"))
only = sys.argv[1:] or list(pkg.CONFIGS)
for cfg in only:
    scene, W, H, _, depth = pkg.CONFIGS[cfg]
    w, h = W // DIV, H // DIV
    dst = os.path.join(ROOT, "tests", "golden", f"{cfg}{SUFFIX}.npz")
    if os.path.exists(dst):
        print(cfg, "exists, skipping", flush=True)
        continue
    spec = pkg.make_scene(scene, 1)
    ob = orc.OracleBackend()
    pkg.scene.emit(spec.world, ob)
    t0 = time.time()
    halves = []
    counters = None
    for seed in (101, 202):
        s, sq, c = ob.render(spec.camera, w, h, SPP // 2, depth, spec.background, seed=seed, threads=threads, tile_size=TILE, want_sumsq=True)
        halves.append((np.nan_to_num(s.astype(np.float64)), np.nan_to_num(sq.astype(np.float64))))
        counters = c
        print(cfg, "half", seed, f"{time.time() - t0:.0f}s", flush=True)
    total = halves[0][0] + halves[1][0]
    total_sq = halves[0][1] + halves[1][1]
    mean = total / SPP
    var = np.maximum(total_sq / SPP - mean * mean, 0.0)
    img = np.sqrt(mean)
    ia, ib = np.sqrt(halves[0][0] / (SPP // 2)), np.sqrt(halves[1][0] / (SPP // 2))
    meta = {"config": cfg, "scene": scene, "width": w, "height": h, "full_width": W, "full_height": H, "spp": SPP,
            "depth": depth, "scene_seed": 1, "seeds": [101, 202], "mae_half_vs_half": float(np.abs(ia - ib).mean()),
            "psnr_half_vs_half": float(10 * np.log10(1.0 / max(1e-12, ((np.clip(ia, 0, 1) - np.clip(ib, 0, 1)) ** 2).mean()))),
            "rays_per_path": counters.rays / counters.paths, "seconds": time.time() - t0, "threads": threads}
    np.savez_compressed(dst, lin=mean.astype(np.float32), sigma=np.sqrt(var / SPP).astype(np.float32), meta=json.dumps(meta))
    print(cfg, json.dumps(meta), flush=True)
