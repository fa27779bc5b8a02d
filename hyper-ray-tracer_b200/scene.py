"""Scene-construction API mirroring the reference's constructors (same names, argument order and meaning).

The reference builds its world from `Box<dyn Hittable>` values with these constructors
(/root/reference/src/hittable/*.rs, src/materials/*.rs, src/textures/*.rs, src/camera.rs:34).  `dyn Hittable`
is opaque, so the B200 path cannot inspect an existing tree; instead the same constructor calls produce light
descriptor objects, and :func:`emit` replays them as C-ABI builder calls (include/hrt.h) on a backend.  The
backend is any object exposing the builder methods — the product backend is
:class:`native.HrtBackend` (libhrt.so); the test-only CPU oracle exposes the same calls, so one description
drives both with identical id allocation.
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass, field
from typing import Any, List as _List, Optional, Sequence, Tuple

import numpy as np

Vec3 = Tuple[float, float, float]


class Plane:  # src/hittable/rect.rs:13-17
    XY = 0
    YZ = 1
    ZX = 2


class Axis:  # src/hittable/rotation.rs:13-17
    X = 0
    Y = 1
    Z = 2


# ---- textures (src/textures/*.rs) --------------------------------------------------------------------
@dataclass(eq=False)
class SolidColor:  # solid_color.rs:15
    color: Vec3


@dataclass(eq=False)
class CheckerTexture:  # checker_texture.rs:16
    odd: Any
    even: Any


@dataclass(eq=False)
class PerlinNoise:
    """Tables of src/perlin_noise.rs:13-18.  The reference fills them from thread_rng (:23-64); here they are
    drawn from an explicit :class:`SceneRng` so that one scene *instance* can be handed to several backends."""

    random_vectors: np.ndarray  # (256, 3) float32
    permutation_x: np.ndarray  # (256,) uint32
    permutation_y: np.ndarray
    permutation_z: np.ndarray

    @staticmethod
    def new(rng: "SceneRng") -> "PerlinNoise":
        vecs = np.zeros((256, 3), dtype=np.float32)
        for i in range(256):  # perlin_noise.rs:27-34
            v = np.array([rng.gen_range(-1.0, 1.0), rng.gen_range(-1.0, 1.0), rng.gen_range(-1.0, 1.0)], dtype=np.float32)
            mag = np.sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2], dtype=np.float32)
            vecs[i] = v * (np.float32(1.0) / mag)  # cgmath normalize
        perms = []
        for _ in range(3):  # perlin_noise.rs:44-64 (gen_range(0..i) is exclusive)
            p = np.arange(256, dtype=np.uint32)
            for i in range(255, 0, -1):
                target = rng.gen_index(i)
                p[i], p[target] = p[target], p[i]
            perms.append(p)
        return PerlinNoise(vecs, perms[0], perms[1], perms[2])


@dataclass(eq=False)
class NoiseTexture:  # noise_texture.rs:16
    scale: float
    noise: PerlinNoise


@dataclass(eq=False)
class ImageTexture:  # image_texture.rs:19-32 — decoded bytes, row-major, `components` bytes per texel
    data: Optional[np.ndarray]  # (h, w, c) uint8 or None (empty)

    @staticmethod
    def empty() -> "ImageTexture":
        return ImageTexture(None)


# ---- materials (src/materials/*.rs) ------------------------------------------------------------------
@dataclass(eq=False)
class Lambertian:  # lambertian.rs:21
    albedo: Any


@dataclass(eq=False)
class Metal:  # metal.rs:23
    albedo: Vec3
    fuzz: float


@dataclass(eq=False)
class Dielectric:  # dielectric.rs:23
    index_of_refraction: float


@dataclass(eq=False)
class DiffuseLight:  # diffuse_light.rs:15
    emit: Any


# ---- hittables (src/hittable/*.rs) -------------------------------------------------------------------
@dataclass(eq=False)
class Sphere:  # sphere.rs:23
    center: Vec3
    radius: float
    material: Any


@dataclass(eq=False)
class MovingSphere:  # moving_sphere.rs:26
    center_start: Vec3
    center_end: Vec3
    time_start: float
    time_end: float
    radius: float
    material: Any


@dataclass(eq=False)
class Rect:  # rect.rs:31
    plane: int
    a0: float
    a1: float
    b0: float
    b1: float
    k: float
    material: Any


@dataclass(eq=False)
class Cuboid:  # cuboid.rs:30
    box_min: Vec3
    box_max: Vec3
    material: Any


@dataclass(eq=False)
class Translation:  # translation.rs:15
    hittable: Any
    displacement: Vec3


@dataclass(eq=False)
class Rotation:  # rotation.rs:38
    axis: int
    hittable: Any
    angle: float


@dataclass(eq=False)
class ConstantMedium:  # constant_medium.rs:24
    boundary: Any
    density: float
    texture: Any


@dataclass(eq=False)
class List:  # list.rs:14
    objects: Sequence[Any]


@dataclass(eq=False)
class BvhNode:  # bvh_node.rs:27
    objects: Sequence[Any]
    time_start: float
    time_end: float


@dataclass
class Camera:  # camera.rs:34-44
    look_from: Vec3
    look_at: Vec3
    fov: float
    aperture: float
    focus_dist: float = 10.0  # application.rs:206
    time_0: float = 0.0  # application.rs:207
    time_1: float = 1.0  # application.rs:208


@dataclass
class SceneSpec:
    """What `Application::new` assembles per `--scene` (application.rs:132-211)."""

    name: str
    world: Any
    camera: Camera
    background: Vec3


# ---- seeded stand-in for rand::thread_rng (distributions of rand 0.8.5) ------------------------------
class SceneRng:
    def __init__(self, seed: int):
        self._g = np.random.Generator(np.random.PCG64(seed))

    def _u32(self) -> int:
        return int(self._g.integers(0, 2**32, dtype=np.uint64))

    def gen(self) -> np.float32:
        """`rng.gen::<f32>()`: 24-bit uniform in [0, 1)."""
        return np.float32(self._u32() >> 8) * np.float32(1.0 / 16777216.0)

    def gen_range(self, lo: float, hi: float) -> np.float32:
        """`rng.gen_range(lo..hi)` for f32: 23-bit uniform scaled into the half-open range."""
        lo32, hi32 = np.float32(lo), np.float32(hi)
        scale = hi32 - lo32
        while True:
            v = np.float32(self._u32() >> 9) * np.float32(1.0 / 8388608.0)
            res = np.float32(v * scale + lo32)
            if res < hi32:
                return res

    def gen_index(self, n: int) -> int:
        """`rng.gen_range(0..n)` for integers."""
        return int(self._g.integers(0, n))


# ---- emission onto a backend -------------------------------------------------------------------------
def _f3(v) -> Tuple[float, float, float]:
    return (float(v[0]), float(v[1]), float(v[2]))


class Emitter:
    """Replays a description as builder calls.  Memoises by object identity so a material/texture shared by
    several primitives (the reference `clone()`s them) is created once."""

    def __init__(self, backend):
        self.b = backend
        self._tex = {}
        self._mat = {}
        self._obj = {}
        self.object_ids = {}  # id(python object) -> backend hittable id

    def texture(self, t) -> int:
        key = id(t)
        if key in self._tex:
            return self._tex[key]
        b = self.b
        if isinstance(t, SolidColor):
            r = b.tex_solid(_f3(t.color))
        elif isinstance(t, CheckerTexture):
            odd = self.texture(t.odd)
            even = self.texture(t.even)
            r = b.tex_checker(odd, even)
        elif isinstance(t, NoiseTexture):
            n = t.noise
            r = b.tex_noise(float(t.scale), n.random_vectors, n.permutation_x, n.permutation_y, n.permutation_z)
        elif isinstance(t, ImageTexture):
            r = b.tex_image(t.data)
        else:
            raise TypeError(f"not a texture: {t!r}")
        self._tex[key] = r
        return r

    def material(self, m) -> int:
        key = id(m)
        if key in self._mat:
            return self._mat[key]
        b = self.b
        if isinstance(m, Lambertian):
            r = b.mat_lambertian(self.texture(m.albedo))
        elif isinstance(m, Metal):
            r = b.mat_metal(_f3(m.albedo), float(m.fuzz))
        elif isinstance(m, Dielectric):
            r = b.mat_dielectric(float(m.index_of_refraction))
        elif isinstance(m, DiffuseLight):
            r = b.mat_diffuse_light(self.texture(m.emit))
        else:
            raise TypeError(f"not a material: {m!r}")
        self._mat[key] = r
        return r

    def hittable(self, h) -> int:
        key = id(h)
        if key in self._obj:
            return self._obj[key]
        b = self.b
        if isinstance(h, Sphere):
            r = b.sphere(_f3(h.center), float(h.radius), self.material(h.material))
        elif isinstance(h, MovingSphere):
            r = b.moving_sphere(_f3(h.center_start), _f3(h.center_end), float(h.time_start), float(h.time_end),
                                float(h.radius), self.material(h.material))
        elif isinstance(h, Rect):
            r = b.rect(int(h.plane), float(h.a0), float(h.a1), float(h.b0), float(h.b1), float(h.k), self.material(h.material))
        elif isinstance(h, Cuboid):
            r = b.cuboid(_f3(h.box_min), _f3(h.box_max), self.material(h.material))
        elif isinstance(h, Translation):
            r = b.translate(self.hittable(h.hittable), _f3(h.displacement))
        elif isinstance(h, Rotation):
            r = b.rotate(int(h.axis), self.hittable(h.hittable), float(h.angle))
        elif isinstance(h, ConstantMedium):
            boundary = self.hittable(h.boundary)
            r = b.constant_medium(boundary, float(h.density), self.texture(h.texture))
        elif isinstance(h, List):
            ids = [self.hittable(o) for o in h.objects]
            r = b.list(ids)
        elif isinstance(h, BvhNode):
            ids = [self.hittable(o) for o in h.objects]
            r = b.bvh(ids, float(h.time_start), float(h.time_end))
        else:
            raise TypeError(f"not a hittable: {h!r}")
        self._obj[key] = r
        self.object_ids[key] = r
        return r


def emit(world, backend) -> Emitter:
    """Create `world` on `backend` and commit it; returns the emitter (for id look-ups)."""
    e = Emitter(backend)
    root = e.hittable(world)
    backend.commit(root)
    e.root = root
    return e
