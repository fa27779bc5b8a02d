"""CPU check of the FLATTENED op stream's semantics (hyper-ray-tracer_b200/csrc/hrt_types.h): a small numpy interpreter of
the 32-byte records — written from the record table in hrt_types.h, independent of the CUDA code — must find the same
closest hit (t and primitive) as the CPU oracle's recursive `world.hit` on the same rays.  This pins, without a GPU, what
the flattener emits: depth-first order, skip links, tight / loose box flags, ray-space push / pop runs, cuboid side order.
Media are not interpreted (their boundary sub-streams are skipped); rays whose oracle hit is a medium hit are left out.
"""
import numpy as np
import pytest

from conftest import build_both, make_rays

OP_BOX, OP_BOX_LOOSE, OP_SPHERE, OP_MSPHERE = 0x10, 0x11, 0x20, 0x21
OP_RECT_XY, OP_RECT_YZ, OP_RECT_ZX, OP_CUBOID = 0x30, 0x31, 0x32, 0x33
OP_TRANSLATE, OP_ROTATE, OP_POP, OP_MEDIUM, OP_MEDIUM_SPHERE, OP_END = 0x40, 0x41, 0x42, 0x43, 0x44, 0x50
F = np.float32


def _box(o, d, mn, mx, tmin, closest, loose):
    """aabb.rs:20-47 per axis; `loose` = the reference's form (each axis against (tmin, closest) on its own), else the
    three slab intervals intersected."""
    inv = F(1.0) / d
    t0 = (mn - o) * inv
    t1 = (mx - o) * inv
    lo = np.where(inv < 0, t1, t0)
    hi = np.where(inv < 0, t0, t1)
    if loose:
        a = np.where(lo > tmin, lo, tmin)
        b = np.where(hi < closest[:, None], hi, closest[:, None])
        return ~np.any(b <= a, axis=1)
    lo_all = np.maximum(np.maximum(lo[:, 0], lo[:, 1]), np.maximum(lo[:, 2], tmin))
    hi_all = np.minimum(np.minimum(hi[:, 0], hi[:, 1]), np.minimum(hi[:, 2], closest))
    return ~(hi_all <= lo_all)


def _sphere(o, d, c, r, tmin, closest):
    oc = o - c
    a = (d * d).sum(axis=1, dtype=F)
    half_b = (oc * d).sum(axis=1, dtype=F)
    cc = (oc * oc).sum(axis=1, dtype=F) - r * r
    disc = half_b * half_b - a * cc
    ok = ~(disc < 0)
    sq = np.sqrt(np.where(ok, disc, F(0)))
    r1 = (-half_b - sq) / a
    r2 = (-half_b + sq) / a
    bad1 = (r1 < tmin) | (closest < r1)
    bad2 = (r2 < tmin) | (closest < r2)
    t = np.where(bad1, r2, r1)
    return ok & ~(bad1 & bad2), t


def _rect(o, d, ik, ia, ib, a0, a1, b0, b1, k, tmin, closest):
    t = (k - o[:, ik]) / d[:, ik]
    ok = ~((t < tmin) | (t > closest))
    a = o[:, ia] + t * d[:, ia]
    b = o[:, ib] + t * d[:, ib]
    ok &= ~((a < a0) | (a > a1) | (b < b0) | (b > b1))
    return ok, t


def trace_stream(ops, rays):
    """Closest hit of every ray against the op stream; returns (hit mask, t, prim id)."""
    f = ops.view(np.float32)
    n = len(rays)
    o = rays["o"].astype(F).copy()
    d = rays["d"].astype(F).copy()
    time = rays["time"].astype(F)
    tmin = rays["tmin"].astype(F)
    closest = rays["tmax"].astype(F).copy()
    prim = np.full(n, -1, dtype=np.int64)
    pc_of = np.zeros(n, dtype=np.int64)
    saved = []  # ray-space stack: (ray indices, their o, their d) per pushed level
    pc = 0
    with np.errstate(all="ignore"):
        while pc < len(ops):
            idx = np.nonzero(pc_of == pc)[0]
            w7 = int(ops[pc, 7])
            op, payload = w7 & 0xFF, w7 >> 8
            nxt = pc + 1
            if op == OP_END:
                break
            if idx.size:
                oo, dd, cl, tm = o[idx], d[idx], closest[idx], tmin[idx]
                if op in (OP_BOX, OP_BOX_LOOSE):
                    hit = _box(oo, dd, f[pc, 0:3], f[pc, 4:7], tm[:, None] if op == OP_BOX_LOOSE else tm, cl, op == OP_BOX_LOOSE)
                    pc_of[idx] = np.where(hit, pc + 1, payload)
                    pc += 1
                    continue
                if op == OP_SPHERE or op == OP_MSPHERE:
                    c = f[pc, 0:3][None, :]
                    if op == OP_MSPHERE:  # moving_sphere.rs:53-57, the aux record holds c1, t0, t1
                        c1, t0, t1 = f[pc + 1, 0:3], f[pc + 1, 3], f[pc + 1, 4]
                        c = c + ((time[idx] - t0) / (t1 - t0))[:, None] * (c1 - f[pc, 0:3])[None, :]
                        nxt = pc + 2
                    ok, t = _sphere(oo, dd, c, f[pc, 3], tm, cl)
                    pid = int(ops[pc, 5])
                elif op in (OP_RECT_XY, OP_RECT_YZ, OP_RECT_ZX):
                    ik, ia, ib = {OP_RECT_XY: (2, 0, 1), OP_RECT_YZ: (0, 1, 2), OP_RECT_ZX: (1, 2, 0)}[op]
                    ok, t = _rect(oo, dd, ik, ia, ib, f[pc, 0], f[pc, 1], f[pc, 2], f[pc, 3], f[pc, 4], tm, cl)
                    pid = int(ops[pc, 6])
                elif op == OP_CUBOID:  # cuboid.rs:30-96: XY@max.z, XY@min.z, ZX@max.y, ZX@min.y, YZ@max.x, YZ@min.x
                    mn, mx = f[pc, 0:3], f[pc, 4:7]
                    ok = np.zeros(idx.size, dtype=bool)
                    t = cl.copy()
                    for (ik, ia, ib), k in (((2, 0, 1), mx[2]), ((2, 0, 1), mn[2]), ((1, 2, 0), mx[1]), ((1, 2, 0), mn[1]),
                                            ((0, 1, 2), mx[0]), ((0, 1, 2), mn[0])):
                        h, ts = _rect(oo, dd, ik, ia, ib, mn[ia], mx[ia], mn[ib], mx[ib], k, tm, t)
                        t = np.where(h, ts, t)
                        ok |= h
                    pid = payload
                elif op in (OP_TRANSLATE, OP_ROTATE):
                    run = max(1, payload)
                    for q in range(pc, pc + run):  # a run of pushes is entered in one step
                        saved.append((idx, o[idx].copy(), d[idx].copy()))
                        if (int(ops[q, 7]) & 0xFF) == OP_TRANSLATE:  # translation.rs:25-29
                            o[idx] = o[idx] - f[q, 0:3][None, :]
                        else:  # rotation.rs:103-116
                            sn, cs, axis = f[q, 0], f[q, 1], int(ops[q, 2])
                            ia, ib = (axis + 1) % 3, (axis + 2) % 3
                            for v in (o, d):
                                va, vb = v[idx, ia].copy(), v[idx, ib].copy()
                                v[idx, ia] = cs * va + sn * vb
                                v[idx, ib] = -sn * va + cs * vb
                    pc_of[idx] = pc + run
                    pc += 1
                    continue
                elif op == OP_POP:
                    run = max(1, payload)
                    for _ in range(run):
                        sidx, so, sd = saved.pop()
                        assert np.array_equal(sidx, idx)
                        o[sidx], d[sidx] = so, sd
                    pc_of[idx] = pc + run
                    pc += 1
                    continue
                elif op in (OP_MEDIUM, OP_MEDIUM_SPHERE):
                    pc_of[idx] = payload  # skip the boundary sub-stream
                    pc += 1
                    continue
                else:
                    raise AssertionError(f"unknown opcode {op:#x} at {pc}")
                sel = idx[ok]
                closest[sel] = t[ok]
                prim[sel] = pid
                pc_of[idx] = nxt
            pc += 1
    return prim >= 0, closest, prim


def _rays(orc, ob, spec, n=1500, seed=4):
    rng = np.random.default_rng(seed)
    cam = ob.camera_rays(spec.camera, 200, 150, rng.random((n, 5), dtype=np.float32))
    h = ob.trace_hits(cam, rng.random(n, dtype=np.float32))
    hit = h[(h["hit"] == 1) & np.isfinite(h["p"]).all(axis=1)]
    if len(hit) == 0:
        return cam
    i = rng.integers(0, len(hit), n)
    dirs = rng.normal(size=(n, 3)).astype(np.float32)
    dirs[np.abs(dirs) < 1e-3] = 1e-3
    sec = make_rays(orc, hit["p"][i], dirs, time=rng.random(n, dtype=np.float32))
    return np.concatenate([cam, sec])


@pytest.mark.parametrize("name", ["random", "simple-light", "cornell", "cornell-smoke", "final"])
def test_stream_semantics_match_the_oracle(pkg, orc, name):
    spec = pkg.make_scene(name, seed=4)
    gb, ob, _, _ = build_both(pkg, orc, spec.world)
    rays = _rays(orc, ob, spec)
    want = ob.trace_hits(rays, np.full(len(rays), 0.5, dtype=np.float32))
    hit, t, prim = trace_stream(gb.ops(), rays)
    medium = (want["hit"] == 1) & np.all(want["n"] == 0.0, axis=1)  # constant_medium.rs:69: normal (0,0,0)
    keep = ~medium & np.isfinite(want["t"])
    assert keep.sum() > len(rays) // 2
    assert np.array_equal(hit[keep], want["hit"][keep] == 1), name
    m = keep & (want["hit"] == 1)
    assert m.sum() > len(rays) // 10
    assert np.array_equal(prim[m], want["prim_id"][m]), name
    assert np.allclose(t[m], want["t"][m], rtol=1e-5, atol=0), name


@pytest.mark.parametrize("name", ["random", "cornell", "final"])
def test_sah_trees_find_the_same_hits(pkg, orc, name):
    """hrt_scene_set_bvh_builder(HRT_BVH_SAH): BVHs whose leaves are all sound are flattened from surface-area-heuristic
    trees — same closest hits as the oracle's reference trees, fewer box visits; BVHs with an unsound (axis-swapped ZX
    rect) leaf box keep the reference tree, clipping behaviour included (Cornell)."""
    spec = pkg.make_scene(name, seed=4)
    gb_ref, ob, _, _ = build_both(pkg, orc, spec.world)
    gb = pkg.HrtBackend()
    gb.set_bvh_builder(pkg.native.HRT_BVH_SAH)
    pkg.scene.emit(spec.world, gb)
    i_ref, i = gb_ref.info(), gb.info()
    assert (i.n_ops, i.n_box_ops, i.n_prim_ops, i.n_loose_boxes) == (i_ref.n_ops, i_ref.n_box_ops, i_ref.n_prim_ops, i_ref.n_loose_boxes)
    if name == "cornell":  # its one BVH holds the ZX light: untouched
        assert i.n_bvh_rebuilt == 0 and np.array_equal(gb.ops(), gb_ref.ops())
        return
    assert i.n_bvh_rebuilt >= 1 and not np.array_equal(gb.ops(), gb_ref.ops())
    rays = _rays(orc, ob, spec)
    want = ob.trace_hits(rays, np.full(len(rays), 0.5, dtype=np.float32))
    hit, t, prim = trace_stream(gb.ops(), rays)
    medium = (want["hit"] == 1) & np.all(want["n"] == 0.0, axis=1)
    keep = ~medium & np.isfinite(want["t"])
    assert np.array_equal(hit[keep], want["hit"][keep] == 1)
    m = keep & (want["hit"] == 1)
    assert np.allclose(t[m], want["t"][m], rtol=1e-5, atol=0)
    # Same primitive too, except on EXACT ties between coincident surfaces (include/hrt.h): `final`'s ground boxes share
    # faces, and a ray that lands on a shared face gets the later leaf of the stream's order instead of the reference's.
    other = m & (prim != want["prim_id"])
    assert other.sum() <= 0.005 * m.sum() and np.array_equal(t[other], want["t"][other]), (name, int(other.sum()))
    # skip links still point forward and stay inside the stream
    ops = gb.ops()
    boxes = np.where(((ops[:, 7] & 0xFF) == OP_BOX) | ((ops[:, 7] & 0xFF) == OP_BOX_LOOSE))[0]
    assert np.all((ops[boxes, 7] >> 8) > boxes) and np.all((ops[boxes, 7] >> 8) <= len(ops) - 1)


def test_sah_visits_fewer_boxes_on_the_random_scene(pkg, orc):
    """The point of the option: count box records visited per ray by the interpreter on both streams."""
    spec = pkg.make_scene("random", seed=4)
    gb_ref, ob, _, _ = build_both(pkg, orc, spec.world)
    gb = pkg.HrtBackend()
    gb.set_bvh_builder(pkg.native.HRT_BVH_SAH)
    pkg.scene.emit(spec.world, gb)
    rays = _rays(orc, ob, spec, n=600)

    def visits(ops):
        f = ops.view(np.float32)
        pc_of = np.zeros(len(rays), dtype=np.int64)
        total = 0
        o, d = rays["o"].astype(F), rays["d"].astype(F)
        closest = rays["tmax"].astype(F).copy()
        for pc in range(len(ops)):
            idx = np.nonzero(pc_of == pc)[0]
            w7 = int(ops[pc, 7]); op, payload = w7 & 0xFF, w7 >> 8
            if op == OP_END or idx.size == 0:
                continue
            if op in (OP_BOX, OP_BOX_LOOSE):
                total += idx.size
                with np.errstate(all="ignore"):
                    h = _box(o[idx], d[idx], f[pc, 0:3], f[pc, 4:7], rays["tmin"][idx].astype(F), closest[idx], False)
                pc_of[idx] = np.where(h, pc + 1, payload)
            elif op == OP_SPHERE:
                with np.errstate(all="ignore"):
                    ok, ts = _sphere(o[idx], d[idx], f[pc, 0:3][None, :], f[pc, 3], rays["tmin"][idx].astype(F), closest[idx])
                closest[idx[ok]] = ts[ok]
                pc_of[idx] = pc + 1
            elif op == OP_MSPHERE:
                pc_of[idx] = pc + 2  # moving spheres only shrink `closest`; leaving them out counts an upper bound on both
            else:
                pc_of[idx] = pc + 1
        return total / len(rays)

    v_ref, v_sah = visits(gb_ref.ops()), visits(gb.ops())
    assert v_sah < 0.75 * v_ref, (v_ref, v_sah)


def test_sah_builder_edge_cases(pkg, orc):
    """One- and two-leaf BVHs, nested BVHs, an unknown builder id, and the option after commit."""
    S, N = pkg.scene, pkg.native
    m = S.Lambertian(S.SolidColor((0.5, 0.5, 0.5)))
    one = S.BvhNode([S.Sphere((0, 0, -5), 1.0, m)], 0.0, 1.0)
    two = S.BvhNode([S.Sphere((3, 0, -5), 1.0, m), S.Sphere((-3, 0, -5), 1.0, m)], 0.0, 1.0)
    moving = S.BvhNode([S.MovingSphere((0, 3, -5), (0, 4, -5), 0.0, 1.0, 0.5, m), S.Sphere((0, -3, -5), 0.5, m),
                        S.Sphere((0, -6, -9), 0.5, m)], 0.0, 1.0)
    world = S.BvhNode([one, two, moving, S.Sphere((0, -1000, 0), 990.0, m)], 0.0, 1.0)
    gb_ref, ob, _, _ = build_both(pkg, orc, world)
    gb = pkg.HrtBackend()
    with pytest.raises(pkg.HrtError) as ei:
        gb.set_bvh_builder(7)
    assert ei.value.code == -1
    gb.set_bvh_builder(N.HRT_BVH_SAH)
    S.emit(world, gb)
    assert gb.info().n_bvh_rebuilt == 4 and gb.info().n_ops == gb_ref.info().n_ops
    with pytest.raises(pkg.HrtError):
        gb.set_bvh_builder(N.HRT_BVH_REFERENCE)  # immutable after commit, like every builder call
    rng = np.random.default_rng(1)
    d = rng.normal(size=(800, 3)).astype(np.float32)
    d[np.abs(d) < 1e-3] = 1e-3
    rays = make_rays(orc, np.tile(np.float32([0, 0, 4]), (800, 1)), d, time=rng.random(800, dtype=np.float32))
    want = ob.trace_hits(rays, np.full(800, 0.5, dtype=np.float32))
    for ops in (gb_ref.ops(), gb.ops()):
        hit, t, prim = trace_stream(ops, rays)
        assert np.array_equal(hit, want["hit"] == 1) and (want["hit"] == 1).sum() > 100
        k = want["hit"] == 1
        assert np.array_equal(prim[k], want["prim_id"][k]) and np.allclose(t[k], want["t"][k], rtol=1e-5, atol=0)


def test_builder_default_can_come_from_the_environment(pkg, monkeypatch):
    """HRT_BVH_BUILDER=sah (what `bench.py --bvh sah` sets) is the default of scenes created afterwards; an explicit
    hrt_scene_set_bvh_builder still wins."""
    spec = pkg.make_scene("random", seed=2)
    monkeypatch.setenv("HRT_BVH_BUILDER", "sah")
    a = pkg.HrtBackend()
    pkg.scene.emit(spec.world, a)
    b = pkg.HrtBackend()
    b.set_bvh_builder(pkg.native.HRT_BVH_REFERENCE)
    pkg.scene.emit(spec.world, b)
    monkeypatch.delenv("HRT_BVH_BUILDER")
    c = pkg.HrtBackend()
    pkg.scene.emit(spec.world, c)
    assert a.info().n_bvh_rebuilt == 1 and b.info().n_bvh_rebuilt == 0 and c.info().n_bvh_rebuilt == 0
    assert np.array_equal(b.ops(), c.ops()) and not np.array_equal(a.ops(), c.ops())


def test_sah_for_sphere_only_bvhs_keeps_every_primitive(pkg, orc):
    """HRT_BVH_SAH_SPHERES: only BVHs made of distinct (moving) spheres are rebuilt — `final` keeps the reference tree over
    its face-sharing ground boxes, so every hit names the oracle's primitive, ties included; duplicates veto the rebuild."""
    S, N = pkg.scene, pkg.native
    spec = pkg.make_scene("final", seed=4)
    gb_ref, ob, _, _ = build_both(pkg, orc, spec.world)
    gb = pkg.HrtBackend()
    gb.set_bvh_builder(N.HRT_BVH_SAH_SPHERES)
    S.emit(spec.world, gb)
    assert gb.info().n_bvh_rebuilt == 1  # the 1000-sphere cube; not the ground boxes, not the top-level BVH
    rays = _rays(orc, ob, spec)
    want = ob.trace_hits(rays, np.full(len(rays), 0.5, dtype=np.float32))
    hit, t, prim = trace_stream(gb.ops(), rays)
    medium = (want["hit"] == 1) & np.all(want["n"] == 0.0, axis=1)
    keep = ~medium & np.isfinite(want["t"])
    m = keep & (want["hit"] == 1)
    assert np.array_equal(hit[keep], want["hit"][keep] == 1)
    assert np.array_equal(prim[m], want["prim_id"][m]) and np.allclose(t[m], want["t"][m], rtol=1e-5, atol=0)

    r = pkg.make_scene("random", seed=4)
    g2 = pkg.HrtBackend()
    g2.set_bvh_builder(N.HRT_BVH_SAH_SPHERES)
    S.emit(r.world, g2)
    assert g2.info().n_bvh_rebuilt == 1

    mat = S.Lambertian(S.SolidColor((0.5, 0.5, 0.5)))
    dup = S.BvhNode([S.Sphere((0, 0, 0), 1.0, mat), S.MovingSphere((0, 0, 0), (0, 0, 0), 0.0, 1.0, 1.0, mat),
                     S.Sphere((3, 0, 0), 1.0, mat)], 0.0, 1.0)
    g3 = pkg.HrtBackend()
    g3.set_bvh_builder(N.HRT_BVH_SAH_SPHERES)
    S.emit(dup, g3)
    assert g3.info().n_bvh_rebuilt == 0
