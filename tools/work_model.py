"""Operation counts per path for each BASELINE config, from the CPU oracle's counters (test infrastructure).

Two traversals are counted on the same scene instance: the reference's per-axis ("loose") box test and the
intersected ("tight") one — the latter is the minimum work the semantics require on sound boxes and is what
bench.py's roofline uses as ALGORITHMIC work.  Writes profiles/work_model.json.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
orc = graft.load_oracle()
KEYS = ["rays", "aabb_tests", "sphere_tests", "rect_tests", "medium_queries", "noise_evals", "scatters"]
out = {}
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 4
for cfg, (scene, w, h, samples, depth) in pkg.CONFIGS.items():
    spec = pkg.make_scene(scene, 1)
    ob = orc.OracleBackend()
    pkg.scene.emit(spec.world, ob)
    entry = {"scene": scene, "width": w, "height": h, "samples": samples, "depth": depth, "counted_at_spp": spp}
    for mode, name in ((0, "loose"), (1, "tight")):
        t = time.time()
        _, _, c = ob.render(spec.camera, w, h, spp, depth, spec.background, seed=3, aabb_mode=mode)
        entry[f"{name}_per_path"] = {k: getattr(c, k) / c.paths for k in KEYS}
        entry[f"{name}_cpu_mpaths_s"] = c.paths / c.seconds / 1e6
        print(cfg, name, {k: round(v, 2) for k, v in entry[f"{name}_per_path"].items()}, f"{c.paths / c.seconds / 1e6:.3f} Mpaths/s",
              f"{time.time() - t:.1f}s", flush=True)
    out[cfg] = entry
with open(os.path.join(ROOT, "profiles", "work_model.json"), "w") as f:
    json.dump(out, f, indent=1)
