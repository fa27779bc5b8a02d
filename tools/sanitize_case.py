"""Smallest case that exercises every kernel (for compute-sanitizer memcheck under gpurun)."""
import sys

import numpy as np

sys.path.insert(0, ".")
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
N = pkg.native
for name in ("final", "cornell-smoke", "random"):
    spec = pkg.make_scene(name, 1)
    r = pkg.renderer.Renderer(spec, device=0)
    for flag, spp in ((N.HRT_FLAG_WAVEFRONT, 160), (N.HRT_FLAG_UNIFORM, 24), (N.HRT_FLAG_INTERPRETER, 8)):
        img, st = r.render(37, 19, spp, 50, seed=1, flags=flag)
        assert st.paths == 37 * 19 * spp
    rays = np.zeros(100, dtype=N.RAY_DTYPE)
    rays["o"] = np.array(spec.camera.look_from, dtype=np.float32)
    rays["d"] = np.random.default_rng(1).normal(size=(100, 3)).astype(np.float32)
    rays["tmin"] = 0.001
    rays["tmax"] = np.inf
    for f in (0, N.HRT_FLAG_EXACT_MATH, N.HRT_FLAG_UNIFORM, N.HRT_FLAG_EXACT_MATH | N.HRT_FLAG_REFERENCE_TRAVERSAL):
        r.backend.trace_hits(rays, None, flags=f)
    r.backend.render_multi([0], spec.camera, 37, 19, 130, 50, spec.background, seed=2)
print("sanitize case ok")
