"""How many box tests would a better BVH or a better visit order save?  (CPU oracle counters only.)

Runs the oracle — tight box test, a few samples per pixel — on each BASELINE config with its study-only builders
(oracle.cpp bvh_study_mode, selected by ORC_BVH_STUDY in a fresh process each): the reference's longest-axis median
split, nearer-child-first order on that tree, a full-sweep SAH tree, and both.  Writes profiles/bvh_study.json.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import json, sys
sys.path.insert(0, %r)
import __graft_entry__ as graft
pkg = graft.load_package(); orc = graft.load_oracle()
out = {}
for cfg in sys.argv[1:]:
    scene, w, h, samples, depth = pkg.CONFIGS[cfg]
    spec = pkg.make_scene(scene, 1)
    ob = orc.OracleBackend(); pkg.scene.emit(spec.world, ob)
    _, _, c = ob.render(spec.camera, w // 2, h // 2, 4, depth, spec.background, seed=3, aabb_mode=1, tile_size=20)
    out[cfg] = {"aabb_tests_per_path": c.aabb_tests / c.paths, "rays_per_path": c.rays / c.paths,
                "sphere_tests_per_path": c.sphere_tests / c.paths, "rect_tests_per_path": c.rect_tests / c.paths}
print(json.dumps(out))
''' % ROOT

cfgs = sys.argv[1:] or ["C1", "C3", "C5"]
res = {}
for mode in ("reference", "near", "sah", "sah+near"):
    env = dict(os.environ)
    env.pop("ORC_BVH_STUDY", None)
    if mode != "reference":
        env["ORC_BVH_STUDY"] = mode
    line = subprocess.run([sys.executable, "-c", CHILD] + cfgs, env=env, capture_output=True, text=True, check=True).stdout.strip().splitlines()[-1]
    res[mode] = json.loads(line)
    print(mode, {c: round(v["aabb_tests_per_path"], 1) for c, v in res[mode].items()}, flush=True)
ref = res["reference"]
for mode in res:
    for c in cfgs:
        res[mode][c]["box_tests_vs_reference"] = res[mode][c]["aabb_tests_per_path"] / ref[c]["aabb_tests_per_path"]
with open(os.path.join(ROOT, "profiles", "bvh_study.json"), "w") as f:
    json.dump(res, f, indent=1)
