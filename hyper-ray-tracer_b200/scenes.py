"""The reference's eight scene generators (/root/reference/src/application.rs:497-935) and the per-scene camera /
background table (application.rs:132-197), restated on the constructor-mirroring API of :mod:`scene`.

The reference draws geometry and perlin tables from an unseeded `rand::thread_rng()` (application.rs:509,820;
perlin_noise.rs:24,59): no two runs of it render the same `random` / `final` / perlin scene.  Here every
generator takes a seed, so ONE scene instance can be built on libhrt.so and on the CPU oracle alike.
All arithmetic that feeds geometry is done in float32, as in the reference.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, Optional

import numpy as np

from .scene import (Axis, BvhNode, Camera, CheckerTexture, ConstantMedium, Cuboid, Dielectric, DiffuseLight, ImageTexture,
                    Lambertian, Metal, MovingSphere, NoiseTexture, PerlinNoise, Plane, Rect, Rotation, SceneRng, SceneSpec,
                    SolidColor, Sphere, Translation)

f32 = np.float32
_ASSET = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "earthmap_rgb8.npz")


def load_earthmap() -> np.ndarray:
    """Decoded texels of the reference's assets/earthmap.jpg (1024x512 RGB8), frozen as a fixture by
    tools/make_earthmap_fixture.py (PIL/libjpeg-turbo decode; SURVEY.md §8c: decoded bytes are 'parity unpinned'
    w.r.t. the Rust jpeg-decoder crate, at most +-1 LSB)."""
    with np.load(_ASSET) as z:
        return np.ascontiguousarray(z["rgb"])


def _v(x, y, z):
    return (f32(x), f32(y), f32(z))


# application.rs:497-565 — NOTE (Q13): this is the *Next Week* variant (checker ground, moving spheres).
def random_scene(seed: int = 1) -> SceneSpec:
    rng = SceneRng(seed)
    objects = [Sphere(_v(0.0, -1000.0, 0.0), 1000.0,
                      Lambertian(CheckerTexture(SolidColor(_v(0.2, 0.3, 0.1)), SolidColor(_v(0.9, 0.9, 0.9)))))]
    for a in range(-11, 11):
        for b in range(-11, 11):
            choose_material = rng.gen()
            cx = f32(a) + f32(0.9) * rng.gen()
            cz = f32(b) + f32(0.9) * rng.gen()
            center = _v(cx, 0.2, cz)
            d = (center[0] - f32(4.0), center[1] - f32(0.2), center[2] - f32(0.0))
            mag = np.sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2], dtype=f32)
            if mag > f32(0.9):
                if choose_material < f32(0.8):
                    albedo = (rng.gen(), rng.gen(), rng.gen())
                    center_2 = (center[0] + f32(0.0), center[1] + rng.gen_range(0.0, 0.5), center[2] + f32(0.0))
                    objects.append(MovingSphere(center, center_2, 0.0, 1.0, 0.2, Lambertian(SolidColor(albedo))))
                elif choose_material < f32(0.95):
                    albedo = (rng.gen_range(0.5, 1.0), rng.gen_range(0.5, 1.0), rng.gen_range(0.5, 1.0))
                    fuzz = rng.gen_range(0.0, 0.5)
                    objects.append(Sphere(center, 0.2, Metal(albedo, fuzz)))
                else:
                    objects.append(Sphere(center, 0.2, Dielectric(1.5)))
    objects.append(Sphere(_v(0.0, 1.0, 0.0), 1.0, Dielectric(1.5)))
    objects.append(Sphere(_v(-4.0, 1.0, 0.0), 1.0, Lambertian(SolidColor(_v(0.4, 0.2, 0.1)))))
    objects.append(Sphere(_v(4.0, 1.0, 0.0), 1.0, Metal(_v(0.7, 0.6, 0.5), 0.0)))
    world = BvhNode(objects, 0.0, 1.0)
    return SceneSpec("random", world, Camera(_v(13, 2, 3), _v(0, 0, 0), 20.0, 0.1), _v(0.7, 0.8, 1.0))


# application.rs:567-587
def two_spheres(seed: int = 1) -> SceneSpec:
    checker = Lambertian(CheckerTexture(SolidColor(_v(0.2, 0.3, 0.1)), SolidColor(_v(0.9, 0.9, 0.9))))
    objects = [Sphere(_v(0.0, -10.0, 0.0), 10.0, checker), Sphere(_v(0.0, 10.0, 0.0), 10.0, checker)]
    return SceneSpec("two-spheres", BvhNode(objects, 0.0, 1.0), Camera(_v(13, 2, 3), _v(0, 0, 0), 20.0, 0.0), _v(0.7, 0.8, 1.0))


# application.rs:589-602
def two_perlin_spheres(seed: int = 1) -> SceneSpec:
    rng = SceneRng(seed)
    noise = Lambertian(NoiseTexture(4.0, PerlinNoise.new(rng)))
    objects = [Sphere(_v(0.0, -1000.0, 0.0), 1000.0, noise), Sphere(_v(0.0, 2.0, 0.0), 2.0, noise)]
    return SceneSpec("two-perlin-spheres", BvhNode(objects, 0.0, 1.0), Camera(_v(13, 2, 3), _v(0, 0, 0), 20.0, 0.0),
                     _v(0.7, 0.8, 1.0))


# application.rs:604-612
def earth(seed: int = 1, image: Optional[np.ndarray] = None) -> SceneSpec:
    tex = ImageTexture(load_earthmap() if image is None else image)
    objects = [Sphere(_v(0.0, 0.0, 0.0), 2.0, Lambertian(tex))]
    return SceneSpec("earth", BvhNode(objects, 0.0, 1.0), Camera(_v(13, 2, 3), _v(0, 0, 0), 20.0, 0.0), _v(0.7, 0.8, 1.0))


# application.rs:614-637
def simple_light(seed: int = 1) -> SceneSpec:
    rng = SceneRng(seed)
    noise = Lambertian(NoiseTexture(4.0, PerlinNoise.new(rng)))
    objects = [Sphere(_v(0.0, -1000.0, 0.0), 1000.0, noise), Sphere(_v(0.0, 2.0, 0.0), 2.0, noise),
               Rect(Plane.XY, 3.0, 5.0, 1.0, 3.0, -2.0, DiffuseLight(SolidColor(_v(4.0, 4.0, 4.0))))]
    return SceneSpec("simple-light", BvhNode(objects, 0.0, 1.0), Camera(_v(26, 3, 6), _v(0, 2, 0), 20.0, 0.0), _v(0, 0, 0))


def _cornell_walls():
    red = Lambertian(SolidColor(_v(0.65, 0.05, 0.05)))
    white = Lambertian(SolidColor(_v(0.73, 0.73, 0.73)))
    green = Lambertian(SolidColor(_v(0.12, 0.45, 0.15)))
    light = DiffuseLight(SolidColor(_v(15.0, 15.0, 15.0)))
    objects = [
        Rect(Plane.YZ, 0.0, 555.0, 0.0, 555.0, 555.0, green),
        Rect(Plane.YZ, 0.0, 555.0, 0.0, 555.0, 0.0, red),
        Rect(Plane.ZX, 213.0, 343.0, 227.0, 332.0, 554.0, light),
        Rect(Plane.ZX, 0.0, 555.0, 0.0, 555.0, 0.0, white),
        Rect(Plane.ZX, 0.0, 555.0, 0.0, 555.0, 555.0, white),
        Rect(Plane.XY, 0.0, 555.0, 0.0, 555.0, 555.0, white),
    ]
    return objects, white


# application.rs:639-721
def cornell_box(seed: int = 1) -> SceneSpec:
    objects, white = _cornell_walls()
    c1 = Cuboid(_v(0, 0, 0), _v(165.0, 330.0, 165.0), white)
    c1 = Translation(Rotation(Axis.Y, c1, 15.0), _v(265.0, 0.0, 295.0))
    objects.append(c1)
    c2 = Cuboid(_v(0, 0, 0), _v(165.0, 165.0, 165.0), white)
    c2 = Translation(Rotation(Axis.Y, c2, -18.0), _v(130.0, 0.0, 65.0))
    objects.append(c2)
    return SceneSpec("cornell", BvhNode(objects, 0.0, 1.0), Camera(_v(278, 278, -800), _v(278, 278, 0), 40.0, 0.0), _v(0, 0, 0))


# application.rs:723-815
def cornell_smoke(seed: int = 1) -> SceneSpec:
    objects, white = _cornell_walls()
    c1 = Cuboid(_v(0, 0, 0), _v(165.0, 330.0, 165.0), white)
    c1 = Translation(Rotation(Axis.Y, c1, 15.0), _v(265.0, 0.0, 295.0))
    objects.append(ConstantMedium(c1, 0.01, SolidColor(_v(0.0, 0.0, 0.0))))
    c2 = Cuboid(_v(0, 0, 0), _v(165.0, 165.0, 165.0), white)
    c2 = Translation(Rotation(Axis.Y, c2, -18.0), _v(130.0, 0.0, 65.0))
    objects.append(ConstantMedium(c2, 0.01, SolidColor(_v(1.0, 1.0, 1.0))))
    return SceneSpec("cornell-smoke", BvhNode(objects, 0.0, 1.0), Camera(_v(278, 278, -800), _v(278, 278, 0), 40.0, 0.0),
                     _v(0, 0, 0))


# application.rs:817-935 — NOTE (Q14): 20x20 = 400 ground boxes, 1000 small spheres, top level is a BvhNode.
def final_scene(seed: int = 1, image: Optional[np.ndarray] = None) -> SceneSpec:
    rng = SceneRng(seed)
    ground = Lambertian(SolidColor(_v(0.48, 0.83, 0.53)))
    ground_boxes = []
    for i in range(20):
        for j in range(20):
            w = f32(100.0)
            x0 = f32(-1000.0) + f32(i) * w
            z0 = f32(-1000.0) + f32(j) * w
            y0 = f32(0.0)
            x1 = x0 + w
            y1 = rng.gen_range(1.0, 101.0)
            z1 = z0 + w
            ground_boxes.append(Cuboid((x0, y0, z0), (x1, y1, z1), ground))
    objects = [BvhNode(ground_boxes, 0.0, 1.0)]
    objects.append(Rect(Plane.ZX, 123.0, 423.0, 147.0, 412.0, 554.0, DiffuseLight(SolidColor(_v(7.0, 7.0, 7.0)))))
    center_1 = _v(400.0, 400.0, 200.0)
    center_2 = (center_1[0] + f32(30.0), center_1[1] + f32(0.0), center_1[2] + f32(0.0))
    objects.append(MovingSphere(center_1, center_2, 0.0, 1.0, 50.0, Lambertian(SolidColor(_v(0.7, 0.3, 0.1)))))
    objects.append(Sphere(_v(260.0, 150.0, 45.0), 50.0, Dielectric(1.5)))
    objects.append(Sphere(_v(0.0, 150.0, 145.0), 50.0, Metal(_v(0.8, 0.8, 0.9), 1.0)))
    glass = Dielectric(1.5)
    objects.append(Sphere(_v(360.0, 150.0, 145.0), 70.0, glass))
    objects.append(ConstantMedium(Sphere(_v(360.0, 150.0, 145.0), 70.0, glass), 0.2, SolidColor(_v(0.2, 0.4, 0.9))))
    objects.append(ConstantMedium(Sphere(_v(0.0, 0.0, 0.0), 5000.0, Dielectric(1.5)), 0.0001, SolidColor(_v(1.0, 1.0, 1.0))))
    earth_tex = ImageTexture(load_earthmap() if image is None else image)
    objects.append(Sphere(_v(400.0, 200.0, 400.0), 100.0, Lambertian(earth_tex)))
    objects.append(Sphere(_v(220.0, 280.0, 300.0), 80.0, Lambertian(NoiseTexture(0.1, PerlinNoise.new(rng)))))
    white = Lambertian(SolidColor(_v(0.73, 0.73, 0.73)))
    sphere_box = []
    for _ in range(1000):
        c = (rng.gen_range(0.0, 165.0), rng.gen_range(0.0, 165.0), rng.gen_range(0.0, 165.0))
        sphere_box.append(Sphere(c, 10.0, white))
    objects.append(Translation(Rotation(Axis.Y, BvhNode(sphere_box, 0.0, 1.0), 15.0), _v(-100.0, 270.0, 395.0)))
    return SceneSpec("final", BvhNode(objects, 0.0, 1.0), Camera(_v(478, 278, -600), _v(278, 278, 0), 40.0, 0.0), _v(0, 0, 0))


# `--scene` values of src/arguments.rs:10-19 (clap kebab-case)
SCENES: Dict[str, Callable[..., SceneSpec]] = {
    "random": random_scene,
    "two-spheres": two_spheres,
    "two-perlin-spheres": two_perlin_spheres,
    "earth": earth,
    "simple-light": simple_light,
    "cornell": cornell_box,
    "cornell-smoke": cornell_smoke,
    "final": final_scene,
}


def make_scene(name: str, seed: int = 1) -> SceneSpec:
    if name not in SCENES:
        raise KeyError(f"unknown scene '{name}' (expected one of {sorted(SCENES)})")
    return SCENES[name](seed)


# BASELINE.json configs: name -> (scene, width, height, samples, depth)
CONFIGS = {
    "C1": ("random", 400, 225, 100, 50),
    "C2a": ("two-perlin-spheres", 800, 450, 1024, 50),
    "C2b": ("earth", 800, 450, 1024, 50),
    "C3": ("cornell", 600, 600, 4096, 50),
    "C4": ("cornell-smoke", 600, 600, 4096, 50),
    "C5": ("final", 800, 800, 10000, 50),
}
