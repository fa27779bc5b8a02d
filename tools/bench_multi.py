"""Single-process multi-GPU throughput (hrt_render_multi) on a BASELINE config: one JSON line per device count."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
cfg = sys.argv[1] if len(sys.argv) > 1 else "C5"
scene, w, h, spp, depth = pkg.CONFIGS[cfg]
if len(sys.argv) > 2:
    spp = int(sys.argv[2])
spec = pkg.make_scene(scene, 1)
r = pkg.renderer.Renderer(spec, device=0)
n_all = pkg.native.device_count()
out = np.empty((h, w, 4), dtype=np.float32)
for n in [k for k in (1, 2, 4, 8) if k <= n_all]:
    devs = list(range(n))
    r.backend.render_multi(devs, spec.camera, w, h, max(64, spp // 8), depth, spec.background, seed=1, out=out)  # warm-up
    t0 = time.perf_counter()
    _, st = r.backend.render_multi(devs, spec.camera, w, h, spp, depth, spec.background, seed=2, out=out)
    dt = time.perf_counter() - t0
    print(json.dumps({"config": cfg, "n_gpus": n, "api": "hrt_render_multi (single process, host RGBA out)", "spp": spp,
                      "mpaths_per_s_e2e": w * h * spp / dt / 1e6, "wall_ms": dt * 1e3, "kernel_ms_max": st.kernel_ms,
                      "resolve_ms": st.resolve_ms, "d2h_ms": st.d2h_ms, "rays": st.rays}), flush=True)
