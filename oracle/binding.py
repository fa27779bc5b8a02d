"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE (see oracle.cpp header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
`OracleBackend` exposes the same builder-method names as the product's `native.HrtBackend`, so one scene
description (hyper-ray-tracer_b200/scene.py) can be emitted onto both with identical id allocation.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")


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
                ("background", C.c_float * 3), ("tile_size", C.c_int32), ("seed", C.c_uint64), ("threads", C.c_int32),
                ("aabb_mode", C.c_int32), ("verbose", C.c_int32)]


class Counters(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("rays", C.c_uint64), ("aabb_tests", C.c_uint64), ("sphere_tests", C.c_uint64),
                ("rect_tests", C.c_uint64), ("medium_queries", C.c_uint64), ("noise_evals", C.c_uint64),
                ("scatters", C.c_uint64), ("seconds", C.c_double)]


RAY_DTYPE = np.dtype([("o", np.float32, 3), ("d", np.float32, 3), ("time", np.float32), ("tmin", np.float32),
                      ("tmax", np.float32)])
HIT_DTYPE = np.dtype([("hit", np.int32), ("t", np.float32), ("p", np.float32, 3), ("n", np.float32, 3), ("u", np.float32),
                      ("v", np.float32), ("front_face", np.int32), ("material_id", np.int32), ("prim_id", np.int32),
                      ("face", np.int32)])
SCATTER_DTYPE = np.dtype([("scattered", np.int32), ("attenuation", np.float32, 3), ("o", np.float32, 3),
                          ("d", np.float32, 3), ("time", np.float32), ("emitted", np.float32, 3)])

_lib = None


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (building the checker is not using it)."""
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "oracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def load_library() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    build()
    lib = C.CDLL(LIB_PATH)
    vp, i32, f3 = C.c_void_p, C.c_int32, C.POINTER(C.c_float)
    lib.orc_scene_create.restype = vp
    lib.orc_scene_destroy.argtypes = [vp]
    lib.orc_scene_destroy.restype = None
    lib.orc_tex_solid.argtypes = [vp, f3]
    lib.orc_tex_checker.argtypes = [vp, i32, i32]
    lib.orc_tex_noise.argtypes = [vp, C.c_float, f3, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.orc_tex_image.argtypes = [vp, C.POINTER(C.c_uint8), C.c_uint32, C.c_uint32, C.c_uint32]
    lib.orc_mat_lambertian.argtypes = [vp, i32]
    lib.orc_mat_metal.argtypes = [vp, f3, C.c_float]
    lib.orc_mat_dielectric.argtypes = [vp, C.c_float]
    lib.orc_mat_diffuse_light.argtypes = [vp, i32]
    lib.orc_sphere.argtypes = [vp, f3, C.c_float, i32]
    lib.orc_moving_sphere.argtypes = [vp, f3, f3, C.c_float, C.c_float, C.c_float, i32]
    lib.orc_rect.argtypes = [vp, i32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, i32]
    lib.orc_cuboid.argtypes = [vp, f3, f3, i32]
    lib.orc_translate.argtypes = [vp, i32, f3]
    lib.orc_rotate.argtypes = [vp, i32, i32, C.c_float]
    lib.orc_constant_medium.argtypes = [vp, i32, C.c_float, i32]
    lib.orc_list.argtypes = [vp, C.POINTER(i32), i32]
    lib.orc_bvh.argtypes = [vp, C.POINTER(i32), i32, C.c_float, C.c_float]
    lib.orc_scene_commit.argtypes = [vp, i32]
    lib.orc_scene_count.argtypes = [vp]
    lib.orc_scene_count.restype = C.c_uint32
    lib.orc_bvh_leaf_order.argtypes = [vp, i32, C.POINTER(i32), i32]
    lib.orc_bvh_node_count.argtypes = [vp, i32]
    lib.orc_bounding_box.argtypes = [vp, i32, f3]
    lib.orc_camera_init.argtypes = [C.POINTER(CameraDesc), C.POINTER(CameraState)]
    lib.orc_camera_init.restype = None
    lib.orc_camera_rays.argtypes = [C.POINTER(CameraDesc), vp, i32, vp]
    lib.orc_camera_rays.restype = None
    lib.orc_trace_hits.argtypes = [vp, vp, i32, vp, vp, i32]
    lib.orc_tex_value.argtypes = [vp, i32, vp, i32, vp]
    lib.orc_scatter_direct.argtypes = [vp, vp, vp, vp, i32, vp]
    lib.orc_sample.argtypes = [i32, i32, C.c_uint64, i32, vp]
    lib.orc_sample.restype = None
    lib.orc_render.argtypes = [vp, C.POINTER(CameraDesc), C.POINTER(RenderDesc), vp, vp, C.POINTER(Counters)]
    lib.orc_resolve.argtypes = [vp, i32, i32, vp]
    lib.orc_resolve.restype = None
    _lib = lib
    return lib


def _arr3(v):
    return (C.c_float * 3)(float(v[0]), float(v[1]), float(v[2]))


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def camera_desc(cam, width: int, height: int) -> CameraDesc:
    return CameraDesc(_arr3(cam.look_from), _arr3(cam.look_at), float(cam.fov), float(cam.aperture), float(cam.focus_dist),
                      float(cam.time_0), float(cam.time_1), int(width), int(height))


class OracleError(RuntimeError):
    pass


class OracleBackend:
    prefix = "orc_"

    def __init__(self):
        self.lib = load_library()
        self.handle = C.c_void_p(self.lib.orc_scene_create())

    def close(self):
        if getattr(self, "handle", None):
            self.lib.orc_scene_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _check(rc: int) -> int:
        if rc < 0:
            raise OracleError(f"oracle call failed ({rc})")
        return rc

    def tex_solid(self, rgb):
        return self._check(self.lib.orc_tex_solid(self.handle, _arr3(rgb)))

    def tex_checker(self, odd, even):
        return self._check(self.lib.orc_tex_checker(self.handle, odd, even))

    def tex_noise(self, scale, ranvec, px, py, pz):
        rv = np.ascontiguousarray(ranvec, dtype=np.float32)
        a, b, c = (np.ascontiguousarray(p, dtype=np.uint32) for p in (px, py, pz))
        u32p = C.POINTER(C.c_uint32)
        return self._check(self.lib.orc_tex_noise(self.handle, float(scale), rv.ctypes.data_as(C.POINTER(C.c_float)),
                                                  a.ctypes.data_as(u32p), b.ctypes.data_as(u32p), c.ctypes.data_as(u32p)))

    def tex_image(self, data):
        if data is None:
            return self._check(self.lib.orc_tex_image(self.handle, None, 0, 0, 0))
        d = np.ascontiguousarray(data, dtype=np.uint8)
        h, w, comps = d.shape
        return self._check(self.lib.orc_tex_image(self.handle, d.ctypes.data_as(C.POINTER(C.c_uint8)), w, h, comps))

    def mat_lambertian(self, tex):
        return self._check(self.lib.orc_mat_lambertian(self.handle, tex))

    def mat_metal(self, rgb, fuzz):
        return self._check(self.lib.orc_mat_metal(self.handle, _arr3(rgb), float(fuzz)))

    def mat_dielectric(self, ior):
        return self._check(self.lib.orc_mat_dielectric(self.handle, float(ior)))

    def mat_diffuse_light(self, tex):
        return self._check(self.lib.orc_mat_diffuse_light(self.handle, tex))

    def sphere(self, c, r, mat):
        return self._check(self.lib.orc_sphere(self.handle, _arr3(c), float(r), mat))

    def moving_sphere(self, c0, c1, t0, t1, r, mat):
        return self._check(self.lib.orc_moving_sphere(self.handle, _arr3(c0), _arr3(c1), float(t0), float(t1), float(r), mat))

    def rect(self, plane, a0, a1, b0, b1, k, mat):
        return self._check(self.lib.orc_rect(self.handle, plane, a0, a1, b0, b1, k, mat))

    def cuboid(self, mn, mx, mat):
        return self._check(self.lib.orc_cuboid(self.handle, _arr3(mn), _arr3(mx), mat))

    def translate(self, child, d):
        return self._check(self.lib.orc_translate(self.handle, child, _arr3(d)))

    def rotate(self, axis, child, deg):
        return self._check(self.lib.orc_rotate(self.handle, axis, child, float(deg)))

    def constant_medium(self, boundary, density, tex):
        return self._check(self.lib.orc_constant_medium(self.handle, boundary, float(density), tex))

    def list(self, ids: Sequence[int]):
        arr = (C.c_int32 * max(1, len(ids)))(*ids)
        return self._check(self.lib.orc_list(self.handle, arr, len(ids)))

    def bvh(self, ids: Sequence[int], t0, t1):
        arr = (C.c_int32 * max(1, len(ids)))(*ids)
        return self._check(self.lib.orc_bvh(self.handle, arr, len(ids), float(t0), float(t1)))

    def commit(self, root):
        return self._check(self.lib.orc_scene_commit(self.handle, root))

    def count(self) -> int:
        return int(self.lib.orc_scene_count(self.handle))

    def bvh_leaf_order(self, bvh: int):
        n = self._check(self.lib.orc_bvh_leaf_order(self.handle, bvh, None, 0))
        out = (C.c_int32 * max(1, n))()
        self._check(self.lib.orc_bvh_leaf_order(self.handle, bvh, out, n))
        return [int(out[i]) for i in range(n)]

    def bvh_node_count(self, bvh: int) -> int:
        return self._check(self.lib.orc_bvh_node_count(self.handle, bvh))

    def bounding_box(self, obj: int) -> np.ndarray:
        out = (C.c_float * 6)()
        self._check(self.lib.orc_bounding_box(self.handle, obj, out))
        return np.array(list(out), dtype=np.float32)

    def camera_init(self, cd: CameraDesc) -> CameraState:
        st = CameraState()
        self.lib.orc_camera_init(C.byref(cd), C.byref(st))
        return st

    def camera_rays(self, cam, width, height, stuuu) -> np.ndarray:
        cd = camera_desc(cam, width, height)
        stuuu = np.ascontiguousarray(stuuu, dtype=np.float32).reshape(-1, 5)
        out = np.zeros(stuuu.shape[0], dtype=RAY_DTYPE)
        self.lib.orc_camera_rays(C.byref(cd), _ptr(stuuu), stuuu.shape[0], _ptr(out))
        return out

    def trace_hits(self, rays: np.ndarray, xi: Optional[np.ndarray] = None, aabb_mode: int = 0) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        n = rays.shape[0]
        out = np.zeros(n, dtype=HIT_DTYPE)
        xi_c = None if xi is None else np.ascontiguousarray(xi, dtype=np.float32)
        self._check(self.lib.orc_trace_hits(self.handle, _ptr(rays), n, None if xi_c is None else _ptr(xi_c), _ptr(out),
                                            aabb_mode))
        return out

    def tex_value(self, tex: int, uvp: np.ndarray) -> np.ndarray:
        uvp = np.ascontiguousarray(uvp, dtype=np.float32).reshape(-1, 5)
        out = np.zeros((uvp.shape[0], 3), dtype=np.float32)
        self._check(self.lib.orc_tex_value(self.handle, tex, _ptr(uvp), uvp.shape[0], _ptr(out)))
        return out

    def scatter(self, rays, hits, u4) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.ascontiguousarray(hits, dtype=HIT_DTYPE)
        u4 = np.ascontiguousarray(u4, dtype=np.float32).reshape(-1, 4)
        out = np.zeros(rays.shape[0], dtype=SCATTER_DTYPE)
        self._check(self.lib.orc_scatter_direct(self.handle, _ptr(rays), _ptr(hits), _ptr(u4), rays.shape[0], _ptr(out)))
        return out

    def render(self, cam, width, height, samples, depth, background, seed=1, threads=0, tile_size=80, aabb_mode=0,
               want_sumsq=False, verbose=False):
        """Application::render on the CPU.  Returns (sum_rgb[h,w,3] f32 linear, sumsq or None, Counters)."""
        cd = camera_desc(cam, width, height)
        rd = RenderDesc(int(width), int(height), int(samples), int(depth), _arr3(background), int(tile_size), int(seed),
                        int(threads), int(aabb_mode), 1 if verbose else 0)
        s = np.zeros((height, width, 3), dtype=np.float32)
        sq = np.zeros((height, width, 3), dtype=np.float32) if want_sumsq else None
        cnt = Counters()
        self._check(self.lib.orc_render(self.handle, C.byref(cd), C.byref(rd), _ptr(s), None if sq is None else _ptr(sq),
                                        C.byref(cnt)))
        return s, sq, cnt


def resolve(sum_rgb: np.ndarray, samples: int) -> np.ndarray:
    """The reference's gamma resolve (application.rs:451-456) -> (h, w, 4) float32."""
    lib = load_library()
    s = np.ascontiguousarray(sum_rgb, dtype=np.float32)
    h, w, _ = s.shape
    out = np.zeros((h, w, 4), dtype=np.float32)
    lib.orc_resolve(_ptr(s), h * w, int(samples), _ptr(out))
    return out


def sample(kind: int, mode: int, seed: int, n: int) -> np.ndarray:
    lib = load_library()
    out = np.zeros((n, 3), dtype=np.float32)
    lib.orc_sample(kind, mode, seed, n, _ptr(out))
    return out
