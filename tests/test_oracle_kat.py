"""Pin the CPU oracle (oracle/oracle.cpp) against known-answer vectors authored from the reference source.

The reference ships no tests, golden vectors or fixtures, and cannot be built here (no Rust
toolchain), so every vector below is derived by hand from the cited reference lines."""
import math

import numpy as np
import pytest

from conftest import make_rays

INF = np.inf


def _hit(ob, orc, o, d, time=0.0, tmin=0.001, tmax=INF, xi=None, aabb_mode=0):
    rays = make_rays(orc, [o], [d], time, tmin, tmax)
    return ob.trace_hits(rays, None if xi is None else np.array([xi], dtype=np.float32), aabb_mode)[0]


def _single(pkg, orc, obj):
    ob = orc.OracleBackend()
    e = pkg.scene.emit(obj, ob)
    return ob, e


# ---- (1) BVH topology fixtures: src/hittable/bvh_node.rs:27-63 ------------------------------------------
def test_cornell_topology(pkg, orc):
    spec = pkg.make_scene("cornell", 1)
    ob, e = _single(pkg, orc, spec.world)
    labels = ["green", "red", "light", "floor", "ceiling", "back", "box1", "box2"]  # push order, application.rs:647-718
    names = {e.object_ids[id(o)]: l for o, l in zip(spec.world.objects, labels)}
    order = [names[i] for i in ob.bvh_leaf_order(e.root)]
    assert order == ["box2", "floor", "red", "ceiling", "back", "light", "box1", "green"]
    assert ob.bvh_node_count(e.root) == 15
    assert ob.count() == 8  # Rotation::count() == 1 (rotation.rs:140-142)


def test_final_topology(pkg, orc):
    spec = pkg.make_scene("final", 1)
    ob, e = _single(pkg, orc, spec.world)
    labels = ["groundBVH", "light", "moving", "glass", "metal", "boundary70", "medium70", "fog5000", "earth", "noise",
              "sphereBox"]
    names = {e.object_ids[id(o)]: l for o, l in zip(spec.world.objects, labels)}
    order = [names[i] for i in ob.bvh_leaf_order(e.root)]
    assert order == ["groundBVH", "metal", "fog5000", "sphereBox", "noise", "glass", "boundary70", "medium70", "earth",
                     "light", "moving"]
    assert ob.bvh_node_count(e.root) == 21
    ground = e.object_ids[id(spec.world.objects[0])]
    assert ob.bvh_node_count(ground) == 799
    sphere_bvh = e.object_ids[id(spec.world.objects[10].hittable.hittable)]
    assert ob.bvh_node_count(sphere_bvh) == 1999
    # 400 cuboids x 6 rects + 9 prims + Rotation counts 1 (rotation.rs:140)
    assert ob.count() == 2400 + 9 + 1


def test_bvh_split_rule(pkg, orc):
    """Largest-extent axis, centroid-sum sort, left = objs[..n/2] (bvh_node.rs:28-52)."""
    S = pkg.scene
    m = S.Lambertian(S.SolidColor((0.5, 0.5, 0.5)))
    # extents: x 0..21, y 0..2, z 0..2 -> axis x; centroids 1, 10, 20 -> left = [a], right = [b, c]
    a = S.Sphere((1, 1, 1), 1.0, m)
    b = S.Sphere((10, 1, 1), 1.0, m)
    c = S.Sphere((20, 1, 1), 1.0, m)
    world = S.BvhNode([c, a, b], 0.0, 1.0)
    ob, e = _single(pkg, orc, world)
    ids = [e.object_ids[id(x)] for x in (a, b, c)]
    assert ob.bvh_leaf_order(e.root) == ids
    assert ob.bvh_node_count(e.root) == 5
    # ties keep insertion order (insertion sort for n <= 20)
    t1 = S.Sphere((0, 0, 0), 1.0, m)
    t2 = S.Sphere((0, 5, 0), 1.0, m)  # y extent largest -> axis y, no tie
    t3 = S.Sphere((0, 5, 0), 1.0, m)  # same centroid as t2
    ob2, e2 = _single(pkg, orc, S.BvhNode([t3, t2, t1], 0.0, 1.0))
    assert ob2.bvh_leaf_order(e2.root) == [e2.object_ids[id(t1)], e2.object_ids[id(t3)], e2.object_ids[id(t2)]]


# ---- (2) analytic hit records per primitive -------------------------------------------------------------
def test_sphere_hit_record(pkg, orc):
    """src/hittable/sphere.rs:40-75: unit sphere at origin, ray from (0,0,-3) along +z."""
    S = pkg.scene
    ob, e = _single(pkg, orc, S.Sphere((0, 0, 0), 1.0, S.Dielectric(1.5)))
    h = _hit(ob, orc, (0, 0, -3), (0, 0, 1))
    assert h["hit"] == 1 and h["t"] == 2.0
    assert tuple(h["p"]) == (0.0, 0.0, -1.0) and tuple(h["n"]) == (0.0, 0.0, -1.0) and h["front_face"] == 1
    # calculate_uv (:31-36): theta = acos(-y) = pi/2 -> v = .5 ; phi = atan2(-z, x) + pi = atan2(1, 0) + pi = 1.5 pi -> u = .75
    assert h["u"] == pytest.approx(0.75, abs=1e-7) and h["v"] == pytest.approx(0.5, abs=1e-7)
    # from inside: far root, front_face false, normal flipped
    h = _hit(ob, orc, (0, 0, 0), (0, 2, 0))
    assert h["hit"] == 1 and h["t"] == 0.5 and tuple(h["n"]) == (0.0, -1.0, 0.0) and h["front_face"] == 0
    # direction is not normalised: t scales
    h = _hit(ob, orc, (0, 0, -3), (0, 0, 4))
    assert h["t"] == 0.5
    # miss
    assert _hit(ob, orc, (0, 2, -3), (0, 0, 1))["hit"] == 0
    # range: tmax below the near root and the far root -> None; tmin past near root -> far root
    assert _hit(ob, orc, (0, 0, -3), (0, 0, 1), tmax=1.5)["hit"] == 0
    assert _hit(ob, orc, (0, 0, -3), (0, 0, 1), tmin=2.5)["t"] == 4.0


def test_moving_sphere(pkg, orc):
    """src/hittable/moving_sphere.rs:53-96: centre interpolates linearly in ray.time."""
    S = pkg.scene
    ob, e = _single(pkg, orc, S.MovingSphere((0, 0, 0), (0, 2, 0), 0.0, 1.0, 1.0, S.Dielectric(1.5)))
    assert _hit(ob, orc, (0, 0, -3), (0, 0, 1), time=0.0)["t"] == 2.0
    h = _hit(ob, orc, (0, 0, -3), (0, 0, 1), time=0.5)  # centre (0,1,0): grazing at y=0 -> disc 0 -> t = 3
    assert h["hit"] == 1 and h["t"] == 3.0
    assert _hit(ob, orc, (0, 0, -3), (0, 0, 1), time=1.0)["hit"] == 0
    bb = ob.bounding_box(e.root)
    assert tuple(bb) == (-1.0, -1.0, -1.0, 1.0, 3.0, 1.0)  # union of boxes at t0 and t1 (:98-110)


def test_rect_axes_and_uv(pkg, orc):
    """src/hittable/rect.rs:53-86.  ZX maps (k=y, a=z, b=x): first range is z, second is x (Q12)."""
    S = pkg.scene
    m = S.Dielectric(1.5)
    ob, _ = _single(pkg, orc, S.Rect(S.Plane.XY, 0.0, 2.0, 0.0, 4.0, 1.0, m))
    h = _hit(ob, orc, (0.5, 1.0, -1.0), (0, 0, 2))
    assert h["hit"] == 1 and h["t"] == 1.0 and h["u"] == 0.25 and h["v"] == 0.25
    assert tuple(h["n"]) == (0.0, 0.0, -1.0) and h["front_face"] == 0  # d.n_out = +2 -> back face, normal flipped
    ob, _ = _single(pkg, orc, S.Rect(S.Plane.YZ, 0.0, 2.0, 0.0, 4.0, 1.0, m))
    h = _hit(ob, orc, (3.0, 0.5, 1.0), (-1, 0, 0))
    assert h["hit"] == 1 and h["t"] == 2.0 and h["u"] == 0.25 and h["v"] == 0.25 and tuple(h["n"]) == (1.0, 0.0, 0.0)
    assert h["front_face"] == 1
    ob, e = _single(pkg, orc, S.Rect(S.Plane.ZX, 0.0, 2.0, 0.0, 4.0, 1.0, m))  # z in [0,2], x in [0,4]
    h = _hit(ob, orc, (3.0, 3.0, 0.5), (0, -1, 0))
    assert h["hit"] == 1 and h["t"] == 2.0 and h["u"] == 0.25 and h["v"] == 0.75 and tuple(h["n"]) == (0.0, 1.0, 0.0)
    assert _hit(ob, orc, (0.5, 3.0, 3.0), (0, -1, 0))["hit"] == 0  # z = 3 is outside the z-range [0,2]
    # bounding box is built x in [a0,a1], z in [b0,b1]  (rect.rs:98-101) — axis-swapped (Q2)
    bb = ob.bounding_box(e.root)
    assert bb[0] == 0.0 and bb[3] == 2.0 and bb[2] == 0.0 and bb[5] == 4.0
    assert bb[1] == np.float32(1.0) - np.float32(0.0001) and bb[4] == np.float32(1.0) + np.float32(0.0001)
    # inclusive edges (a < a0 || a > a1 rejects)
    ob, _ = _single(pkg, orc, S.Rect(S.Plane.XY, 0.0, 2.0, 0.0, 4.0, 1.0, m))
    assert _hit(ob, orc, (2.0, 4.0, 0.0), (0, 0, 1))["hit"] == 1


def test_cuboid_sides(pkg, orc):
    """src/hittable/cuboid.rs:30-96 side order; list.rs:20-31 closest-so-far."""
    S = pkg.scene
    ob, e = _single(pkg, orc, S.Cuboid((0, 0, 0), (1, 2, 3), S.Dielectric(1.5)))
    cases = [((0.5, 1, 5), (0, 0, -1), 0, 2.0), ((0.5, 1, -2), (0, 0, 1), 1, 2.0), ((0.5, 4, 1), (0, -1, 0), 2, 2.0),
             ((0.5, -2, 1), (0, 1, 0), 3, 2.0), ((3, 1, 1), (-1, 0, 0), 4, 2.0), ((-2, 1, 1), (1, 0, 0), 5, 2.0)]
    for o, d, face, t in cases:
        h = _hit(ob, orc, o, d)
        # every Rect's outward normal is +k (rect.rs:81-83): the three min-side faces report front_face = false
        # when hit from outside the box
        assert h["hit"] == 1 and h["face"] == face and h["t"] == t and h["front_face"] == (1 if face % 2 == 0 else 0), (o, d, h)
    assert ob.count() == 6
    h = _hit(ob, orc, (0.5, 1, 1), (0, 0, 1))  # from inside: exit face, normal flipped inward
    assert h["face"] == 0 and h["t"] == 2.0 and h["front_face"] == 0 and tuple(h["n"]) == (0.0, 0.0, -1.0)


def test_translation_forces_front_face(pkg, orc):
    """src/hittable/translation.rs:24-37 — Q4: set_face_normal is re-applied to the already-flipped normal."""
    S = pkg.scene
    inner = S.Sphere((0, 0, 0), 1.0, S.Dielectric(1.5))
    ob, _ = _single(pkg, orc, S.Translation(inner, (10, 0, 0)))
    h = _hit(ob, orc, (10, 0, -3), (0, 0, 1))
    assert h["t"] == 2.0 and tuple(h["p"]) == (10.0, 0.0, -1.0) and h["front_face"] == 1
    h = _hit(ob, orc, (10, 0, 0), (0, 0, 1))  # from INSIDE: plain sphere reports front_face = false ...
    assert h["t"] == 1.0 and tuple(h["n"]) == (0.0, 0.0, -1.0)
    assert h["front_face"] == 1  # ... Translation turns it into true (Q4)


def test_rotation_y(pkg, orc):
    """src/hittable/rotation.rs:102-134 with Axis::Y -> (r,a,b) = (1,2,0)."""
    S = pkg.scene
    rect = S.Rect(S.Plane.YZ, -1.0, 1.0, -1.0, 1.0, 2.0, S.Dielectric(1.5))  # plane x = 2, normal +x
    ob, e = _single(pkg, orc, S.Rotation(S.Axis.Y, rect, 90.0))
    # object->world: z' = cos z - sin x ; x' = sin z + cos x.  With 90deg the plane x=2 maps to z = -2.
    h = _hit(ob, orc, (0.25, 0.5, -5.0), (0, 0, 1))
    assert h["hit"] == 1 and h["t"] == pytest.approx(3.0, abs=1e-5)
    assert h["p"][2] == pytest.approx(-2.0, abs=1e-5) and h["p"][0] == pytest.approx(0.25, abs=1e-5)
    assert h["n"][2] == pytest.approx(-1.0, abs=1e-6) and abs(h["n"][0]) < 1e-6
    assert h["front_face"] == 1
    assert ob.count() == 1


def test_constant_medium(pkg, orc):
    """src/hittable/constant_medium.rs:34-76 with an injected draw."""
    S = pkg.scene
    boundary = S.Sphere((0, 0, 0), 1.0, S.Dielectric(1.5))
    ob, e = _single(pkg, orc, S.ConstantMedium(boundary, 0.5, S.SolidColor((1, 1, 1))))
    xi = 0.5
    h = _hit(ob, orc, (0, 0, -3), (0, 0, 1), xi=xi)
    hit_distance = np.float32(-1.0 / 0.5) * (np.log(np.float32(xi), dtype=np.float32) / np.log(np.float32(math.e), dtype=np.float32))
    assert h["hit"] == 1 and h["t"] == pytest.approx(2.0 + float(hit_distance), rel=1e-6)
    assert tuple(h["n"]) == (0.0, 0.0, 0.0) and h["front_face"] == 0 and h["u"] == 0 and h["v"] == 0
    assert h["material_id"] == 1  # the Isotropic allocated after the boundary's Dielectric
    # too small a draw -> distance beyond the chord -> None
    assert _hit(ob, orc, (0, 0, -3), (0, 0, 1), xi=0.1)["hit"] == 0
    # ray starting inside: r1.t clamped to tmin then to 0
    h = _hit(ob, orc, (0, 0, 0), (0, 0, 1), xi=0.9)
    assert h["hit"] == 1 and h["t"] == pytest.approx(0.001 + float(np.float32(-2.0) * np.log(np.float32(0.9))), rel=1e-5)
    # direction length matters: distance = (t2 - t1) * |d|
    h2 = _hit(ob, orc, (0, 0, -3), (0, 0, 2), xi=xi)
    assert h2["t"] == pytest.approx(1.0 + float(hit_distance) / 2.0, rel=1e-6)


# ---- (3) Q1/Q2: loose per-axis AABB + axis-swapped ZX box + visit order clip the Cornell light ------------
def test_cornell_light_clipping_vectors(pkg, orc):
    spec = pkg.make_scene("cornell", 1)
    ob, e = _single(pkg, orc, spec.world)
    ids = {l: e.object_ids[id(o)] for o, l in zip(spec.world.objects,
                                                  ["green", "red", "light", "floor", "ceiling", "back", "box1", "box2"])}
    o = np.array([278, 278, -800], dtype=np.float32)
    targets = [(300, 554, 220), (300, 554, 280), (300, 554, 340), (220, 554, 280), (340, 554, 280)]
    rays = make_rays(orc, [o] * 5, [np.array(t, dtype=np.float32) - o for t in targets])
    h = ob.trace_hits(rays)
    expect = ["ceiling", "light", "light", "ceiling", "ceiling"]
    assert [int(x) for x in h["prim_id"]] == [ids[n] for n in expect]
    assert h["t"][1] == 1.0 and h["t"][2] == 1.0
    assert h["t"][0] == pytest.approx(1.0036232, rel=1e-6)
    # a correct (intersected) slab test returns the ceiling for the third vector: it discriminates Q1
    h_tight = ob.trace_hits(rays, aabb_mode=1)
    assert int(h_tight["prim_id"][2]) == ids["ceiling"]


# ---- (4) perlin KAT on a fixed table (checks Q6) -----------------------------------------------------------
def _perlin_reference(ranvec, px, py, pz, p):
    """Direct numpy restatement of src/perlin_noise.rs:80-123 in float32."""
    f = np.float32
    p = np.asarray(p, dtype=f)
    fl = np.floor(p)
    i, j, k = int(fl[0]), int(fl[1]), int(fl[2])
    u, v, w = p - fl
    u = u * u * (f(3) - f(2) * u)
    v = v * v * (f(3) - f(2) * v)
    w = w * w * (f(3) - f(2) * w)
    acc = f(0)
    for x in range(2):
        for y in range(2):
            for z in range(2):
                g = ranvec[int(px[(i + x) & 255]) ^ int(py[(j + y) & 255]) ^ int(pz[(k + z) & 255])]
                wt = np.array([u - f(x), v - f(y), w - f(z)], dtype=f)
                d = f(f(g[0] * wt[0]) + f(g[1] * wt[1])) + f(g[2] * wt[2])
                acc = f(acc + f(f(f((f(x) * u + f(1 - x) * (f(1) - u)) * (f(y) * v + f(1 - y) * (f(1) - v))) *
                                    (f(z) * w + f(1 - z) * (f(1) - w))) * d))
    return acc


def test_perlin_kat(pkg, orc):
    S = pkg.scene
    noise = S.PerlinNoise.new(S.SceneRng(123))
    tex = S.NoiseTexture(4.0, noise)
    ob = orc.OracleBackend()
    em = S.Emitter(ob)
    tid = em.texture(tex)
    pts = np.array([[0.3, 1.7, -2.2], [10.5, -0.25, 3.125], [-7.75, 2.5, 0.0625]], dtype=np.float32)
    uvp = np.concatenate([np.zeros((3, 2), np.float32), pts], axis=1)
    got = ob.tex_value(tid, uvp)
    for q, g in zip(pts, got):
        # noise_texture.rs:25-31 (Q7: turbulence is fed scale*p), perlin_noise.rs:66-78
        sp = np.float32(4.0) * q
        acc, wgt, pp = np.float32(0), np.float32(1), sp.copy()
        for _ in range(7):
            acc = np.float32(acc + wgt * _perlin_reference(noise.random_vectors, noise.permutation_x, noise.permutation_y,
                                                           noise.permutation_z, pp))
            wgt = np.float32(wgt * np.float32(0.5))
            pp = pp * np.float32(2.0)
        val = np.float32(0.5) * (np.float32(1.0) + np.sin(np.float32(sp[2] + np.float32(10.0) * np.abs(acc)), dtype=np.float32))
        assert g[0] == g[1] == g[2]
        assert float(g[0]) == pytest.approx(float(val), abs=2e-6)
    # tables: unit vectors and true permutations (perlin_noise.rs:27-64)
    assert np.allclose(np.linalg.norm(noise.random_vectors, axis=1), 1.0, atol=1e-6)
    for p in (noise.permutation_x, noise.permutation_y, noise.permutation_z):
        assert sorted(p.tolist()) == list(range(256))


# ---- (5) image-texel KATs ---------------------------------------------------------------------------------
def test_image_texel_kat(pkg, orc):
    """src/textures/image_texture.rs:36-63: v is flipped, nearest texel, clamp to w-1/h-1, NaN -> texel 0."""
    S = pkg.scene
    img = np.zeros((2, 4, 3), dtype=np.uint8)  # h=2, w=4
    for j in range(2):
        for i in range(4):
            img[j, i] = (10 * i, 100 * j + 5, 255 if (i + j) % 2 else 0)
    ob = orc.OracleBackend()
    tid = S.Emitter(ob).texture(S.ImageTexture(img))
    nan = np.nan
    uvp = np.array([[0, 0, 0, 0, 0], [0.5, 0.5, 0, 0, 0], [1, 1, 0, 0, 0], [nan, nan, 0, 0, 0], [-3, 7, 0, 0, 0],
                    [0.26, 0.49, 9, 9, 9]], dtype=np.float32)
    got = ob.tex_value(tid, uvp)
    sc = np.float32(1.0 / 255.0)

    def texel(i, j):
        return tuple(sc * np.float32(c) for c in img[j, i])

    assert tuple(got[0]) == texel(0, 1)  # u=0 -> i=0 ; v=0 -> 1-0=1 -> j=2 -> clamped to 1
    assert tuple(got[1]) == texel(2, 1)  # u=.5 -> i=2 ; v=.5 -> .5*2 = 1 -> j=1
    assert tuple(got[2]) == texel(3, 0)  # u=1 -> i=4 -> 3 ; v=1 -> 0
    assert tuple(got[3]) == texel(0, 0)  # NaN clamps to NaN; `as u32` of NaN = 0 (both i and j)
    assert tuple(got[4]) == texel(0, 0)  # u<0 -> 0 ; v>1 -> 1 -> 1-1 = 0
    assert tuple(got[5]) == texel(1, 1)  # u=.26 -> 1.04 -> 1 ; v=.49 -> .51*2 = 1.02 -> 1
    empty = S.Emitter(ob).texture(S.ImageTexture.empty())
    assert tuple(ob.tex_value(empty, uvp[:1])[0]) == (1.0, 0.0, 1.0)  # :37-39


def test_checker_kat(pkg, orc):
    """src/textures/checker_texture.rs:22-30: sign of sin(10x) sin(10y) sin(10z); negative -> odd."""
    S = pkg.scene
    ob = orc.OracleBackend()
    tid = S.Emitter(ob).texture(S.CheckerTexture(S.SolidColor((1, 0, 0)), S.SolidColor((0, 1, 0))))
    pts = np.array([[0.1, 0.1, 0.1], [0.1, 0.1, -0.1], [0.4, 0.1, 0.1], [0.0, 0.3, 0.3]], dtype=np.float32)
    uvp = np.concatenate([np.zeros((4, 2), np.float32), pts], axis=1)
    got = ob.tex_value(tid, uvp)
    assert tuple(got[0]) == (0, 1, 0)  # +++ -> even
    assert tuple(got[1]) == (1, 0, 0)  # ++- -> odd
    assert tuple(got[2]) == (1, 0, 0)  # sin(4) < 0
    assert tuple(got[3]) == (0, 1, 0)  # product == 0 is not < 0 -> even


# ---- camera (src/camera.rs:34-95) -------------------------------------------------------------------------
def test_camera_kat(pkg, orc):
    S = pkg.scene
    cam = S.Camera((0, 0, 0), (0, 0, -1), 90.0, 0.0, focus_dist=1.0)
    ob = orc.OracleBackend()
    st = ob.camera_init(orc.camera_desc(cam, 200, 100))
    # w = (0,0,1); u = (0,1,0) x w = (1,0,0); v = w x u = (0,1,0); viewport 4 x 2 at focus 1
    assert tuple(st.w) == (0.0, 0.0, 1.0) and tuple(st.u) == (1.0, 0.0, 0.0) and tuple(st.v) == (0.0, 1.0, 0.0)
    assert st.horizontal[0] == pytest.approx(4.0, rel=1e-6) and st.vertical[1] == pytest.approx(2.0, rel=1e-6)
    assert tuple(st.lower_left_corner) == pytest.approx((-2.0, -1.0, -1.0), rel=1e-6)
    r = ob.camera_rays(cam, 200, 100, np.array([[0.5, 0.5, 0.3, 0.7, 0.25]], dtype=np.float32))[0]
    assert tuple(r["o"]) == (0.0, 0.0, 0.0) and tuple(r["d"]) == pytest.approx((0.0, 0.0, -1.0), abs=1e-6)
    assert r["time"] == 0.25
    # defocus: lens_radius = aperture / 2 (camera.rs:57); offset = u*rd.x + v*rd.y
    cam2 = S.Camera((0, 0, 0), (0, 0, -1), 90.0, 2.0, focus_dist=1.0)
    r = ob.camera_rays(cam2, 200, 100, np.array([[0.5, 0.5, 1.0, 0.0, 0.0]], dtype=np.float32))[0]
    assert tuple(r["o"]) == pytest.approx((1.0, 0.0, 0.0), abs=1e-6)  # disk sample r=1, phi=0
    assert tuple(r["d"]) == pytest.approx((-1.0, 0.0, -1.0), abs=1e-6)


# ---- math.rs ----------------------------------------------------------------------------------------------
def test_scatter_kats(pkg, orc):
    """reflect / refract / schlick (math.rs:47-62) via Metal and Dielectric with injected uniforms."""
    S = pkg.scene
    metal = S.Metal((0.8, 0.6, 0.4), 0.0)
    glass = S.Dielectric(1.5)
    world = S.List([S.Sphere((0, 0, 0), 1.0, metal), S.Sphere((10, 0, 0), 1.0, glass)])
    ob, e = _single(pkg, orc, world)
    rays = make_rays(orc, [(0, 2, -2), (10, 0, -3), (10, 0, 0)], [(0, -1, 1), (0, 0, 1), (0, 0, 1)])
    hits = ob.trace_hits(rays)
    assert list(hits["hit"]) == [1, 1, 1]
    u4 = np.array([[0.1, 0.2, 0.3, 0.4], [0.99, 0, 0, 0], [0.99, 0, 0, 0]], dtype=np.float32)
    out = ob.scatter(rays, hits, u4)
    # metal: 45deg onto the top of the sphere? hit point (0, 1, -1)/.. check reflect(v,n) = v - 2(v.n)n with unit v
    n = hits["n"][0]
    v = np.array([0, -1, 1], np.float32) / np.sqrt(np.float32(2))
    expect = v - 2 * np.dot(v, n) * n
    assert out["scattered"][0] == 1 and np.allclose(out["d"][0], expect, atol=1e-6)
    assert tuple(out["attenuation"][0]) == pytest.approx((0.8, 0.6, 0.4))
    # dielectric, normal incidence, xi=.99 > schlick(1, 1/1.5)=0.04 -> refract straight through
    assert out["scattered"][1] == 1 and np.allclose(out["d"][1], (0, 0, 1), atol=1e-6)
    assert tuple(out["attenuation"][1]) == (1.0, 1.0, 1.0)
    # from inside (front_face false -> ratio 1.5), normal incidence -> straight through
    assert np.allclose(out["d"][2], (0, 0, 1), atol=1e-6)
    # xi below the reflectance -> reflect
    out2 = ob.scatter(rays[1:2], hits[1:2], np.array([[0.01, 0, 0, 0]], np.float32))
    assert np.allclose(out2["d"][0], (0, 0, -1), atol=1e-6)


def test_samplers_match_in_distribution(orc):
    """The fixed-draw samplers (used by the CUDA path) sample the SAME distributions as the reference's rejection
    loops (math.rs:12-40): compare low-order moments and radial CDFs."""
    n = 200_000
    for kind in (0, 1, 2):
        a = orc.sample(kind, 0, 11, n).astype(np.float64)
        b = orc.sample(kind, 1, 12, n).astype(np.float64)
        assert np.allclose(a.mean(0), b.mean(0), atol=0.01)
        assert np.allclose((a * a).mean(0), (b * b).mean(0), atol=0.01)
        ra, rb = np.linalg.norm(a, axis=1), np.linalg.norm(b, axis=1)
        qs = np.linspace(0.05, 0.95, 19)
        assert np.allclose(np.quantile(ra, qs), np.quantile(rb, qs), atol=0.01)
        if kind == 1:
            assert np.allclose(ra, 1.0, atol=1e-5) and np.allclose(rb, 1.0, atol=1e-5)
        if kind == 2:
            assert np.all(a[:, 2] == 0) and np.all(b[:, 2] == 0)


def test_render_conventions(pkg, orc):
    """application.rs:437-456: bottom-up rows, /(w-1), /(h-1), sqrt(sum/spp); background-only scene is exact."""
    S = pkg.scene
    far = S.Sphere((0, 0, 1000), 1.0, S.Dielectric(1.5))  # behind the camera
    ob, e = _single(pkg, orc, S.BvhNode([far], 0.0, 1.0))
    cam = S.Camera((0, 0, 0), (0, 0, -1), 40.0, 0.0)
    s, sq, cnt = ob.render(cam, 16, 8, 4, 5, (0.25, 0.5, 1.0), seed=3, threads=2, tile_size=5, want_sumsq=True)
    assert s.shape == (8, 16, 3) and np.all(s == np.array([1.0, 2.0, 4.0], np.float32))
    assert cnt.paths == 16 * 8 * 4 and cnt.rays == cnt.paths
    img = orc.resolve(s, 4)
    assert np.all(img[..., 3] == 1.0) and np.allclose(img[..., :3], np.sqrt([0.25, 0.5, 1.0]))
    # depth 0 -> black (application.rs:478-480)
    s0, _, c0 = ob.render(cam, 4, 4, 2, 0, (1, 1, 1), seed=1, threads=1)
    assert np.all(s0 == 0) and c0.rays == 0
