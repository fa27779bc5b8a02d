"""ctypes binding of libhrt.so (include/hrt.h).  There is no CPU fallback: if the CUDA library is missing the
import of this module raises, and every compute call raises :class:`HrtError` when no B200 is present."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HRT_LIB") or os.path.join(_HERE, "csrc", "libhrt.so")  # HRT_LIB: diagnostic builds

HRT_FLAG_REFERENCE_TRAVERSAL = 1
HRT_FLAG_EXACT_MATH = 2
HRT_FLAG_INTERPRETER = 8
HRT_FLAG_UNIFORM = 64
HRT_FLAG_WAVEFRONT = 128
ABI_VERSION = 3  # include/hrt.h HRT_ABI_VERSION
HRT_BVH_REFERENCE = 0
HRT_BVH_TREES = 1
HRT_STREAM_REFERENCE = 0
HRT_STREAM_FAST = 1
HRT_STREAM_WAVE = 2


PROGRESS_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p)


def _check_frame(out, width, height):
    """A caller-provided frame buffer goes to C as a raw pointer: it must be exactly (height, width, 4) float32, C order."""
    if out is None:
        return np.empty((height, width, 4), dtype=np.float32)
    if not (isinstance(out, np.ndarray) and out.dtype == np.float32 and out.shape == (height, width, 4) and out.flags.c_contiguous
            and out.flags.writeable):
        raise ValueError(f"out must be a writable C-contiguous float32 array of shape ({height}, {width}, 4)")
    return out


class HrtError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libhrt error {code}: {message}")
        self.code = code
        self.message = message


class CameraDesc(C.Structure):
    _fields_ = [("look_from", C.c_float * 3), ("look_at", C.c_float * 3), ("vfov", C.c_float), ("aperture", C.c_float),
                ("focus_dist", C.c_float), ("time0", C.c_float), ("time1", C.c_float), ("width", C.c_int32),
                ("height", C.c_int32)]


class CameraState(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("lower_left_corner", C.c_float * 3), ("horizontal", C.c_float * 3),
                ("vertical", C.c_float * 3), ("u", C.c_float * 3), ("v", C.c_float * 3), ("w", C.c_float * 3),
                ("lens_radius", C.c_float), ("time0", C.c_float), ("time1", C.c_float)]


class RenderDesc(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("samples", C.c_int32), ("depth", C.c_int32),
                ("background", C.c_float * 3), ("seed", C.c_uint64), ("sample_begin", C.c_int32),
                ("sample_count", C.c_int32), ("flags", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("rays", C.c_uint64), ("kernel_ms", C.c_float), ("resolve_ms", C.c_float),
                ("h2d_ms", C.c_float), ("d2h_ms", C.c_float), ("launches", C.c_int32), ("grid", C.c_int32),
                ("block", C.c_int32)]


class SceneView(C.Structure):
    _fields_ = [("look_from", C.c_float * 3), ("look_at", C.c_float * 3), ("vfov", C.c_float), ("aperture", C.c_float),
                ("focus_dist", C.c_float), ("time0", C.c_float), ("time1", C.c_float), ("background", C.c_float * 3)]


class Peaks(C.Structure):
    _fields_ = [("fp32_tflops", C.c_float), ("l2_read_gbs", C.c_float), ("fma_ms", C.c_float), ("l2_ms", C.c_float),
                ("sm_count", C.c_int32), ("clock_khz", C.c_int32)]


class SceneInfo(C.Structure):
    _fields_ = [("n_ops", C.c_int32), ("n_box_ops", C.c_int32), ("n_loose_boxes", C.c_int32), ("n_prim_ops", C.c_int32),
                ("n_materials", C.c_int32), ("n_textures", C.c_int32), ("n_noise_tables", C.c_int32),
                ("n_images", C.c_int32), ("n_media", C.c_int32), ("n_contexts", C.c_int32),
                ("max_context_depth", C.c_int32), ("time_min", C.c_float), ("time_max", C.c_float),
                ("n_fast_ops", C.c_int32), ("n_fast_box_ops", C.c_int32), ("n_bvh_trees", C.c_int32),
                ("n_tree_nodes", C.c_int32), ("max_tree_depth", C.c_int32)]


RAY_DTYPE = np.dtype([("o", np.float32, 3), ("d", np.float32, 3), ("time", np.float32), ("tmin", np.float32),
                      ("tmax", np.float32)])
HIT_DTYPE = np.dtype([("hit", np.int32), ("t", np.float32), ("p", np.float32, 3), ("n", np.float32, 3), ("u", np.float32),
                      ("v", np.float32), ("front_face", np.int32), ("material_id", np.int32), ("prim_id", np.int32),
                      ("face", np.int32)])
SCATTER_DTYPE = np.dtype([("scattered", np.int32), ("attenuation", np.float32, 3), ("o", np.float32, 3),
                          ("d", np.float32, 3), ("time", np.float32), ("emitted", np.float32, 3)])

# every symbol include/hrt.h declares
EXPORTS = [
    "hrt_last_error", "hrt_abi_version", "hrt_device_count", "hrt_scene_create", "hrt_scene_destroy", "hrt_tex_solid",
    "hrt_tex_checker", "hrt_tex_noise", "hrt_tex_image", "hrt_mat_lambertian", "hrt_mat_metal", "hrt_mat_dielectric",
    "hrt_mat_diffuse_light", "hrt_sphere", "hrt_moving_sphere", "hrt_rect", "hrt_cuboid", "hrt_translate", "hrt_rotate",
    "hrt_constant_medium", "hrt_list", "hrt_bvh", "hrt_scene_commit", "hrt_scene_count", "hrt_scene_get_info",
    "hrt_scene_get_ops", "hrt_bvh_leaf_order", "hrt_bounding_box", "hrt_camera_init", "hrt_scene_upload", "hrt_render",
    "hrt_render_accum", "hrt_render_accum_device", "hrt_resolve_device", "hrt_trace_hits", "hrt_tex_value",
    "hrt_scatter", "hrt_camera_rays", "hrt_philox_uniforms", "hrt_scene_evict", "hrt_scene_device_bytes",
    "hrt_measure_peaks", "hrt_scene_refresh", "hrt_render_multi", "hrt_render_accum_multi",
    "hrt_scene_set_bvh_builder", "hrt_scene_get_tree_nodes", "hrt_scene_get_tree_spans", "hrt_render_progressive", "hrt_make_scene", "hrt_scene_save",
    "hrt_scene_load",
]

_lib = None


def load_library(path: Optional[str] = None) -> C.CDLL:
    """Load libhrt.so and declare its prototypes.  Raises OSError when the CUDA extension has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise OSError(f"{p} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'); "
                      "there is no CPU fallback")
    lib = C.CDLL(p)
    f3 = C.POINTER(C.c_float)
    vp = C.c_void_p
    i32 = C.c_int32
    lib.hrt_last_error.restype = C.c_char_p
    lib.hrt_last_error.argtypes = []
    lib.hrt_abi_version.restype = i32
    if lib.hrt_abi_version() != ABI_VERSION:  # the struct layouts below are those of this ABI version
        raise OSError(f"{p} has ABI version {lib.hrt_abi_version()}, this binding needs {ABI_VERSION}: rebuild the extension")
    lib.hrt_device_count.restype = i32
    lib.hrt_scene_create.argtypes = [C.POINTER(vp)]
    lib.hrt_scene_destroy.argtypes = [vp]
    lib.hrt_scene_destroy.restype = None
    lib.hrt_tex_solid.argtypes = [vp, f3]
    lib.hrt_tex_checker.argtypes = [vp, i32, i32]
    lib.hrt_tex_noise.argtypes = [vp, C.c_float, f3, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.hrt_tex_image.argtypes = [vp, C.POINTER(C.c_uint8), C.c_uint32, C.c_uint32, C.c_uint32]
    lib.hrt_mat_lambertian.argtypes = [vp, i32]
    lib.hrt_mat_metal.argtypes = [vp, f3, C.c_float]
    lib.hrt_mat_dielectric.argtypes = [vp, C.c_float]
    lib.hrt_mat_diffuse_light.argtypes = [vp, i32]
    lib.hrt_sphere.argtypes = [vp, f3, C.c_float, i32]
    lib.hrt_moving_sphere.argtypes = [vp, f3, f3, C.c_float, C.c_float, C.c_float, i32]
    lib.hrt_rect.argtypes = [vp, i32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, i32]
    lib.hrt_cuboid.argtypes = [vp, f3, f3, i32]
    lib.hrt_translate.argtypes = [vp, i32, f3]
    lib.hrt_rotate.argtypes = [vp, i32, i32, C.c_float]
    lib.hrt_constant_medium.argtypes = [vp, i32, C.c_float, i32]
    lib.hrt_list.argtypes = [vp, C.POINTER(i32), i32]
    lib.hrt_bvh.argtypes = [vp, C.POINTER(i32), i32, C.c_float, C.c_float]
    lib.hrt_scene_commit.argtypes = [vp, i32]
    lib.hrt_scene_count.argtypes = [vp]
    lib.hrt_scene_get_info.argtypes = [vp, C.POINTER(SceneInfo)]
    lib.hrt_scene_get_ops.argtypes = [vp, i32, vp, i32]
    lib.hrt_scene_get_tree_nodes.argtypes = [vp, vp, i32]
    lib.hrt_scene_get_tree_spans.argtypes = [vp, C.POINTER(i32), i32]
    lib.hrt_scene_set_bvh_builder.argtypes = [vp, i32]
    lib.hrt_bvh_leaf_order.argtypes = [vp, i32, C.POINTER(i32), i32]
    lib.hrt_bounding_box.argtypes = [vp, i32, f3]
    lib.hrt_camera_init.argtypes = [C.POINTER(CameraDesc), C.POINTER(CameraState)]
    lib.hrt_scene_upload.argtypes = [vp, i32]
    lib.hrt_render.argtypes = [vp, i32, C.POINTER(CameraDesc), C.POINTER(RenderDesc), vp, C.POINTER(Stats)]
    lib.hrt_render_accum.argtypes = [vp, i32, C.POINTER(CameraDesc), C.POINTER(RenderDesc), vp, C.POINTER(Stats)]
    lib.hrt_render_accum_device.argtypes = [vp, i32, C.POINTER(CameraDesc), C.POINTER(RenderDesc), vp, vp, C.POINTER(Stats)]
    lib.hrt_resolve_device.argtypes = [i32, vp, i32, i32, i32, vp, vp]
    lib.hrt_trace_hits.argtypes = [vp, i32, vp, i32, vp, vp, C.c_uint32]
    lib.hrt_tex_value.argtypes = [vp, i32, i32, vp, i32, vp, C.c_uint32]
    lib.hrt_scatter.argtypes = [vp, i32, vp, vp, vp, i32, vp, C.c_uint32]
    lib.hrt_camera_rays.argtypes = [i32, C.POINTER(CameraDesc), vp, i32, vp, C.c_uint32]
    lib.hrt_philox_uniforms.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, f3]
    lib.hrt_scene_evict.argtypes = [vp, i32]
    lib.hrt_make_scene.argtypes = [vp, C.c_char_p, C.c_uint64, C.POINTER(C.c_uint8), C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(i32),
                                   C.POINTER(SceneView)]
    lib.hrt_scene_save.argtypes = [vp, i32, C.POINTER(SceneView), C.c_char_p]
    lib.hrt_scene_load.argtypes = [C.c_char_p, C.POINTER(vp), C.POINTER(i32), C.POINTER(SceneView)]
    lib.hrt_render_progressive.argtypes = [vp, i32, C.POINTER(CameraDesc), C.POINTER(RenderDesc), i32, PROGRESS_FN, vp, vp, C.POINTER(Stats)]
    lib.hrt_render_multi.argtypes = [vp, C.POINTER(i32), i32, C.POINTER(CameraDesc), C.POINTER(RenderDesc), vp, C.POINTER(Stats)]
    lib.hrt_render_accum_multi.argtypes = [vp, C.POINTER(i32), i32, C.POINTER(CameraDesc), C.POINTER(RenderDesc), vp, C.POINTER(Stats)]
    lib.hrt_scene_refresh.argtypes = [vp, i32]
    lib.hrt_scene_device_bytes.argtypes = [vp]
    lib.hrt_measure_peaks.argtypes = [i32, C.POINTER(Peaks)]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("hrt_last_error", "hrt_scene_destroy", "hrt_scene_device_bytes"):
            fn.restype = i32
    lib.hrt_scene_device_bytes.restype = C.c_int64
    if path is None:
        _lib = lib
    return lib


def _arr3(v):
    return (C.c_float * 3)(float(v[0]), float(v[1]), float(v[2]))


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def camera_desc(cam, width: int, height: int) -> CameraDesc:
    return CameraDesc(_arr3(cam.look_from), _arr3(cam.look_at), float(cam.fov), float(cam.aperture), float(cam.focus_dist),
                      float(cam.time_0), float(cam.time_1), int(width), int(height))


class HrtBackend:
    """Builder + compute calls on one `hrt_scene` handle.  Method names are the hrt.h names without prefix."""

    prefix = "hrt_"

    def __init__(self, lib: Optional[C.CDLL] = None):
        self.lib = lib or load_library()
        h = C.c_void_p()
        self._check(self.lib.hrt_scene_create(C.byref(h)))
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            self.lib.hrt_scene_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int) -> int:
        if rc < 0:
            raise HrtError(rc, (self.lib.hrt_last_error() or b"").decode("utf-8", "replace"))
        return rc

    # ---- builder ----
    def tex_solid(self, rgb):
        return self._check(self.lib.hrt_tex_solid(self.handle, _arr3(rgb)))

    def tex_checker(self, odd, even):
        return self._check(self.lib.hrt_tex_checker(self.handle, odd, even))

    def tex_noise(self, scale, ranvec, px, py, pz):
        rv = np.ascontiguousarray(ranvec, dtype=np.float32)
        a, b, c = (np.ascontiguousarray(p, dtype=np.uint32) for p in (px, py, pz))
        u32p = C.POINTER(C.c_uint32)
        return self._check(self.lib.hrt_tex_noise(self.handle, float(scale), rv.ctypes.data_as(C.POINTER(C.c_float)),
                                                  a.ctypes.data_as(u32p), b.ctypes.data_as(u32p), c.ctypes.data_as(u32p)))

    def tex_image(self, data):
        if data is None:
            return self._check(self.lib.hrt_tex_image(self.handle, None, 0, 0, 0))
        d = np.ascontiguousarray(data, dtype=np.uint8)
        h, w, comps = d.shape
        return self._check(self.lib.hrt_tex_image(self.handle, d.ctypes.data_as(C.POINTER(C.c_uint8)), w, h, comps))

    def mat_lambertian(self, tex):
        return self._check(self.lib.hrt_mat_lambertian(self.handle, tex))

    def mat_metal(self, rgb, fuzz):
        return self._check(self.lib.hrt_mat_metal(self.handle, _arr3(rgb), float(fuzz)))

    def mat_dielectric(self, ior):
        return self._check(self.lib.hrt_mat_dielectric(self.handle, float(ior)))

    def mat_diffuse_light(self, tex):
        return self._check(self.lib.hrt_mat_diffuse_light(self.handle, tex))

    def sphere(self, c, r, mat):
        return self._check(self.lib.hrt_sphere(self.handle, _arr3(c), float(r), mat))

    def moving_sphere(self, c0, c1, t0, t1, r, mat):
        return self._check(self.lib.hrt_moving_sphere(self.handle, _arr3(c0), _arr3(c1), float(t0), float(t1), float(r), mat))

    def rect(self, plane, a0, a1, b0, b1, k, mat):
        return self._check(self.lib.hrt_rect(self.handle, plane, a0, a1, b0, b1, k, mat))

    def cuboid(self, mn, mx, mat):
        return self._check(self.lib.hrt_cuboid(self.handle, _arr3(mn), _arr3(mx), mat))

    def translate(self, child, d):
        return self._check(self.lib.hrt_translate(self.handle, child, _arr3(d)))

    def rotate(self, axis, child, deg):
        return self._check(self.lib.hrt_rotate(self.handle, axis, child, float(deg)))

    def constant_medium(self, boundary, density, tex):
        return self._check(self.lib.hrt_constant_medium(self.handle, boundary, float(density), tex))

    def list(self, ids: Sequence[int]):
        arr = (C.c_int32 * max(1, len(ids)))(*ids)
        return self._check(self.lib.hrt_list(self.handle, arr, len(ids)))

    def bvh(self, ids: Sequence[int], t0, t1):
        arr = (C.c_int32 * max(1, len(ids)))(*ids)
        return self._check(self.lib.hrt_bvh(self.handle, arr, len(ids), float(t0), float(t1)))

    # ---- the scene library (hrt_make_scene / hrt_scene_save / hrt_scene_load) ----
    def make_scene(self, name: str, seed: int = 1, image: Optional[np.ndarray] = None):
        """The reference's generator `name` (src/arguments.rs:10-19) with an explicit seed, issued on this (uncommitted)
        scene; `image` = decoded texels (h, w, 3|4) uint8 for `earth` / `final`.  Returns (root id, SceneView)."""
        root, view = C.c_int32(-1), SceneView()
        if image is not None:
            d = np.ascontiguousarray(image, dtype=np.uint8)
            h, w, comps = d.shape
            ptr = d.ctypes.data_as(C.POINTER(C.c_uint8))
        else:
            ptr, w, h, comps = None, 0, 0, 0
        self._check(self.lib.hrt_make_scene(self.handle, name.encode(), int(seed), ptr, w, h, comps, C.byref(root), C.byref(view)))
        return int(root.value), view

    def save(self, root: int, view: "SceneView", path: str):
        self._check(self.lib.hrt_scene_save(self.handle, int(root), C.byref(view), path.encode()))

    @classmethod
    def load(cls, path: str):
        """A stored scene instance as a fresh, uncommitted backend: (backend, root id, SceneView)."""
        lib = load_library()
        handle, root, view = C.c_void_p(), C.c_int32(-1), SceneView()
        rc = lib.hrt_scene_load(path.encode(), C.byref(handle), C.byref(root), C.byref(view))
        if rc < 0:
            raise HrtError(rc, lib.hrt_last_error().decode())
        self = cls.__new__(cls)
        self.lib = lib
        self.handle = handle
        return self, int(root.value), view

    def set_bvh_builder(self, builder: int):
        """HRT_BVH_TREES (1, default: sound BVHs become OP_BVH trees in the fast form) or HRT_BVH_REFERENCE (0); before
        commit (include/hrt.h)."""
        return self._check(self.lib.hrt_scene_set_bvh_builder(self.handle, int(builder)))

    def commit(self, root):
        return self._check(self.lib.hrt_scene_commit(self.handle, root))

    # ---- introspection ----
    def count(self) -> int:
        return self._check(self.lib.hrt_scene_count(self.handle))

    def info(self) -> SceneInfo:
        i = SceneInfo()
        self._check(self.lib.hrt_scene_get_info(self.handle, C.byref(i)))
        return i

    def ops(self, which: int = HRT_STREAM_REFERENCE) -> np.ndarray:
        """(n, 8) uint32 records of the reference (default) or the fast form of the op stream (hrt_types.h)."""
        n = self._check(self.lib.hrt_scene_get_ops(self.handle, which, None, 0))
        out = np.zeros((n, 8), dtype=np.uint32)
        self._check(self.lib.hrt_scene_get_ops(self.handle, which, _ptr(out), n))
        return out

    def tree_nodes(self) -> np.ndarray:
        """(n, 16) uint16 view of the fast form's OP_BVH tree nodes (hrt_types.h Bvh2Node): halves 0-5 left box, 6-7 the
        left link (int32), 8-13 right box, 14-15 the right link."""
        n = self._check(self.lib.hrt_scene_get_tree_nodes(self.handle, None, 0))
        out = np.zeros((max(n, 1), 16), dtype=np.uint16)
        self._check(self.lib.hrt_scene_get_tree_nodes(self.handle, _ptr(out), n))
        return out[:n]

    def tree_spans(self) -> np.ndarray:
        """(n, 4) int32 rows {OP_BVH pc, context, from_pc, to_pc} of the fast form's trees outside medium boundaries
        (include/hrt.h hrt_scene_get_tree_spans)."""
        n = self._check(self.lib.hrt_scene_get_tree_spans(self.handle, None, 0))
        out = np.zeros((n, 4), dtype=np.int32)
        if n:
            self._check(self.lib.hrt_scene_get_tree_spans(self.handle, out.ctypes.data_as(C.POINTER(C.c_int32)), n))
        return out

    def bvh_leaf_order(self, bvh: int):
        n = self._check(self.lib.hrt_bvh_leaf_order(self.handle, bvh, None, 0))
        out = (C.c_int32 * max(1, n))()
        self._check(self.lib.hrt_bvh_leaf_order(self.handle, bvh, out, n))
        return [int(out[i]) for i in range(n)]

    def bounding_box(self, obj: int) -> np.ndarray:
        out = (C.c_float * 6)()
        self._check(self.lib.hrt_bounding_box(self.handle, obj, out))
        return np.array(list(out), dtype=np.float32)

    def camera_init(self, cd: CameraDesc) -> CameraState:
        st = CameraState()
        self._check(self.lib.hrt_camera_init(C.byref(cd), C.byref(st)))
        return st

    # ---- compute (CUDA only) ----
    def upload(self, device: int = 0):
        self._check(self.lib.hrt_scene_upload(self.handle, device))

    def evict(self, device: int = 0):
        self._check(self.lib.hrt_scene_evict(self.handle, device))

    def refresh(self, device: int = 0):
        """Re-copy the scene tables host->device into the existing allocations."""
        self._check(self.lib.hrt_scene_refresh(self.handle, device))

    def device_bytes(self) -> int:
        return int(self._check(self.lib.hrt_scene_device_bytes(self.handle)))

    def _render_desc(self, width, height, samples, depth, background, seed, sample_begin, sample_count, flags):
        return RenderDesc(int(width), int(height), int(samples), int(depth), _arr3(background), int(seed), int(sample_begin),
                          int(sample_count), int(flags))

    def render(self, cam, width, height, samples, depth, background, seed=0, device=0, flags=0, resolve=True, out=None):
        """`Application::render` through the C ABI with HOST buffers; returns (h, w, 4) float32, rows bottom-up."""
        cd = camera_desc(cam, width, height)
        rd = self._render_desc(width, height, samples, depth, background, seed, 0, 0, flags)
        out = _check_frame(out, width, height)
        st = Stats()
        fn = self.lib.hrt_render if resolve else self.lib.hrt_render_accum
        self._check(fn(self.handle, device, C.byref(cd), C.byref(rd), _ptr(out), C.byref(st)))
        return out, st

    def render_progressive(self, cam, width, height, samples, depth, background, batch, on_frame=None, seed=0, device=0, flags=0,
                           out=None):
        """hrt_render_progressive: `on_frame(samples_done, samples_total, frame)` after every `batch` samples (frame is a
        view of `out`, valid during the call); a truthy return cancels.  Returns (out, stats, cancelled)."""
        cd = camera_desc(cam, width, height)
        rd = self._render_desc(width, height, samples, depth, background, seed, 0, 0, flags)
        out = _check_frame(out, width, height)
        st = Stats()

        def trampoline(_user, done, total, _ptr_):
            return 1 if (on_frame is not None and on_frame(int(done), int(total), out)) else 0
        cb = PROGRESS_FN(trampoline)
        rc = self.lib.hrt_render_progressive(self.handle, device, C.byref(cd), C.byref(rd), int(batch), cb, None, _ptr(out), C.byref(st))
        if rc < 0:
            self._check(rc)
        return out, st, rc == 1

    def render_multi(self, devices, cam, width, height, samples, depth, background, seed=0, flags=0, resolve=True, out=None):
        """Single-process multi-GPU render (hrt_render_multi): samples sharded over `devices`, fused peer reduce+resolve."""
        cd = camera_desc(cam, width, height)
        rd = self._render_desc(width, height, samples, depth, background, seed, 0, 0, flags)
        out = _check_frame(out, width, height)
        devs = (C.c_int32 * len(devices))(*devices)
        st = Stats()
        fn = self.lib.hrt_render_multi if resolve else self.lib.hrt_render_accum_multi
        self._check(fn(self.handle, devs, len(devices), C.byref(cd), C.byref(rd), _ptr(out), C.byref(st)))
        return out, st

    def render_accum_device(self, cam, width, height, samples, depth, background, seed, device, d_accum_ptr, stream_ptr=0,
                            sample_begin=0, sample_count=0, flags=0, want_stats=False):
        cd = camera_desc(cam, width, height)
        rd = self._render_desc(width, height, samples, depth, background, seed, sample_begin, sample_count, flags)
        st = Stats()
        self._check(self.lib.hrt_render_accum_device(self.handle, device, C.byref(cd), C.byref(rd), C.c_void_p(d_accum_ptr),
                                                     C.c_void_p(stream_ptr), C.byref(st) if want_stats else None))
        return st

    def resolve_device(self, device, d_accum_ptr, width, height, samples, d_out_ptr, stream_ptr=0):
        self._check(self.lib.hrt_resolve_device(device, C.c_void_p(d_accum_ptr), width, height, samples, C.c_void_p(d_out_ptr),
                                                C.c_void_p(stream_ptr)))

    def trace_hits(self, rays: np.ndarray, xi: Optional[np.ndarray] = None, flags: int = 0, device: int = 0) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        n = rays.shape[0]
        out = np.zeros(n, dtype=HIT_DTYPE)
        xi_c = None if xi is None else np.ascontiguousarray(xi, dtype=np.float32)
        self._check(self.lib.hrt_trace_hits(self.handle, device, _ptr(rays), n, None if xi_c is None else _ptr(xi_c), _ptr(out),
                                            flags))
        return out

    def tex_value(self, tex: int, uvp: np.ndarray, flags: int = 0, device: int = 0) -> np.ndarray:
        uvp = np.ascontiguousarray(uvp, dtype=np.float32).reshape(-1, 5)
        out = np.zeros((uvp.shape[0], 3), dtype=np.float32)
        self._check(self.lib.hrt_tex_value(self.handle, device, tex, _ptr(uvp), uvp.shape[0], _ptr(out), flags))
        return out

    def scatter(self, rays, hits, u4, flags: int = 0, device: int = 0) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.ascontiguousarray(hits, dtype=HIT_DTYPE)
        u4 = np.ascontiguousarray(u4, dtype=np.float32).reshape(-1, 4)
        out = np.zeros(rays.shape[0], dtype=SCATTER_DTYPE)
        self._check(self.lib.hrt_scatter(self.handle, device, _ptr(rays), _ptr(hits), _ptr(u4), rays.shape[0], _ptr(out), flags))
        return out

    def camera_rays(self, cam, width, height, stuuu, flags: int = 0, device: int = 0) -> np.ndarray:
        cd = camera_desc(cam, width, height)
        stuuu = np.ascontiguousarray(stuuu, dtype=np.float32).reshape(-1, 5)
        out = np.zeros(stuuu.shape[0], dtype=RAY_DTYPE)
        self._check(self.lib.hrt_camera_rays(device, C.byref(cd), _ptr(stuuu), stuuu.shape[0], _ptr(out), flags))
        return out


def philox_uniforms(seed, pixel, sample, bounce, block):
    lib = load_library()
    out = (C.c_float * 4)()
    lib.hrt_philox_uniforms(int(seed), int(pixel), int(sample), int(bounce), int(block), out)
    return np.array(list(out), dtype=np.float32)


def measure_peaks(device: int = 0) -> Peaks:
    lib = load_library()
    p = Peaks()
    rc = lib.hrt_measure_peaks(device, C.byref(p))
    if rc < 0:
        raise HrtError(rc, (lib.hrt_last_error() or b"").decode())
    return p


def device_count() -> int:
    return int(load_library().hrt_device_count())
