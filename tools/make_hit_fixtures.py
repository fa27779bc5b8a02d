"""The fixed deterministic ray set of the north star, committed: tests/golden/hits_<scene>.npz = rays + medium uniforms +
the CPU oracle's hit records (t, point, normal, front_face, u/v, material / primitive / side ids) for all eight scenes.

Why a file: the parity tests otherwise regenerate their expectation from the live oracle on every run, so an oracle
regression would move the target silently.  tests/test_hit_fixtures.py checks the oracle against the file on the CPU
(bit for bit) and both CUDA builds against it on the GPU.

Ray set per scene (seeded): camera rays through random film points with random lens / time draws, plus secondary rays
leaving the oracle's hit points in random, non-unit directions with the reference's t_min = 0.001 (application.rs:482).
All direction components are non-zero and origins are off every primitive plane.
Regenerate (only when the oracle is deliberately changed): python tools/make_hit_fixtures.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402
from conftest import make_rays  # noqa: E402

SCENES = ["random", "two-spheres", "two-perlin-spheres", "earth", "simple-light", "cornell", "cornell-smoke", "final"]
N_CAM, N_SEC, SCENE_SEED, RAY_SEED, W, H = 2500, 2500, 1, 20261019, 400, 300


def ray_set(orc, spec, ob):
    rng = np.random.default_rng(RAY_SEED)
    stuuu = rng.random((N_CAM, 5), dtype=np.float32)
    cam = ob.camera_rays(spec.camera, W, H, stuuu)
    h = ob.trace_hits(cam, rng.random(N_CAM, dtype=np.float32))
    hit = h[h["hit"] == 1]
    hit = hit[np.isfinite(hit["p"]).all(axis=1)]
    idx = rng.integers(0, len(hit), N_SEC)
    d = rng.normal(size=(N_SEC, 3)).astype(np.float32)
    d[np.abs(d) < 1e-3] = 1e-3
    d *= rng.uniform(0.2, 3.0, size=(N_SEC, 1)).astype(np.float32)
    sec = make_rays(orc, hit["p"][idx], d, time=rng.random(N_SEC, dtype=np.float32))
    rays = np.concatenate([cam, sec])
    xi = rng.random(len(rays), dtype=np.float32)
    return rays, xi


def main():
    pkg, orc = graft.load_package(), graft.load_oracle()
    out_dir = os.path.join(ROOT, "tests", "golden")
    for name in SCENES:
        spec = pkg.make_scene(name, SCENE_SEED)
        ob = orc.OracleBackend()
        pkg.scene.emit(spec.world, ob)
        rays, xi = ray_set(orc, spec, ob)
        hits = ob.trace_hits(rays, xi)
        path = os.path.join(out_dir, f"hits_{name}.npz")
        np.savez_compressed(path, rays=rays, xi=xi, hits=hits, scene=name, scene_seed=SCENE_SEED)
        print(name, len(rays), "rays,", int((hits["hit"] == 1).sum()), "hits,", os.path.getsize(path) // 1024, "KB")


if __name__ == "__main__":
    main()
