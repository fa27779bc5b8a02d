"""CPU check of the FLATTENED op stream's semantics (hyper-ray-tracer_b200/csrc/hrt_types.h): a small numpy interpreter of
the 32-byte records — written from the record table in hrt_types.h, independent of the CUDA code — must find the same
closest hit (t and primitive) as the CPU oracle's recursive `world.hit` on the same rays.  This pins, without a GPU, what
the flattener emits: depth-first order, skip links, tight / loose box flags, ray-space push / pop runs, cuboid side order.
Media are not interpreted (their boundary sub-streams are skipped); rays whose oracle hit is a medium hit are left out.
The FAST form's OP_BVH trees (two-child nodes with fp16 boxes, walked with a stack, nearer child first, an exact tie
going to the later record) are interpreted too, one ray at a time.
"""
import numpy as np
import pytest

from conftest import build_both, make_rays, nested_tree_world

OP_BOX, OP_BOX_LOOSE, OP_SPHERE, OP_MSPHERE = 0x10, 0x11, 0x20, 0x21
OP_RECT_XY, OP_RECT_YZ, OP_RECT_ZX, OP_CUBOID = 0x30, 0x31, 0x32, 0x33
OP_TRANSLATE, OP_ROTATE, OP_POP, OP_MEDIUM, OP_MEDIUM_SPHERE, OP_BVH, OP_MEDIUM_CUBOID, OP_END = 0x40, 0x41, 0x42, 0x43, 0x44, 0x45, 0x46, 0x50
OP_BVH_PRE = 0x47
F = np.float32


def _slab1(o, inv, box, tmin, closest):
    """Entry / exit of ONE ray against a (6,) float32 box; missed iff hi < lo (hrt_device.cuh slab16)."""
    t0 = (box[0:3] - o) * inv
    t1 = (box[3:6] - o) * inv
    lo = np.where(inv < 0, t1, t0)
    hi = np.where(inv < 0, t0, t1)
    return max(np.fmax.reduce(lo), tmin), min(np.fmin.reduce(hi), closest)


def _leaf1(ops, f, pc, o, d, time, tmin, closest):
    """One primitive record against ONE ray: (hit, t, prim id) with the reference's inclusive range checks."""
    one = lambda x: np.asarray([x], dtype=F)
    oo, dd = o[None, :], d[None, :]
    op = int(ops[pc, 7]) & 0xFF
    if op in (OP_SPHERE, OP_MSPHERE):
        c = f[pc, 0:3][None, :]
        if op == OP_MSPHERE:
            c1, t0, t1 = f[pc + 1, 0:3], f[pc + 1, 3], f[pc + 1, 4]
            c = c + ((one(time) - t0) / (t1 - t0))[:, None] * (c1 - f[pc, 0:3])[None, :]
        ok, t = _sphere(oo, dd, c, f[pc, 3], one(tmin), one(closest))
        return bool(ok[0]), t[0], int(ops[pc, 5])
    if op in (OP_RECT_XY, OP_RECT_YZ, OP_RECT_ZX):
        ik, ia, ib = {OP_RECT_XY: (2, 0, 1), OP_RECT_YZ: (0, 1, 2), OP_RECT_ZX: (1, 2, 0)}[op]
        ok, t = _rect(oo, dd, ik, ia, ib, f[pc, 0], f[pc, 1], f[pc, 2], f[pc, 3], f[pc, 4], one(tmin), one(closest))
        return bool(ok[0]), t[0], int(ops[pc, 6])
    assert op == OP_CUBOID, hex(op)
    mn, mx = f[pc, 0:3], f[pc, 4:7]
    hit, t = False, one(closest)
    for (ik, ia, ib), k in (((2, 0, 1), mx[2]), ((2, 0, 1), mn[2]), ((1, 2, 0), mx[1]), ((1, 2, 0), mn[1]),
                            ((0, 1, 2), mx[0]), ((0, 1, 2), mn[0])):
        h, ts = _rect(oo, dd, ik, ia, ib, mn[ia], mx[ia], mn[ib], mx[ib], k, one(tmin), t)
        if h[0]:
            hit, t = True, ts
    return hit, t[0], int(ops[pc, 7]) >> 8


def _ref_leaf_box(ops, f, pc, ts, te):
    """The reference's bounding box of a leaf record (sphere.rs:77-83, moving_sphere.rs:98-110 over the BvhNode's time
    interval, rect.rs:88-103 as written, cuboid.rs:104-106)."""
    op = int(ops[pc, 7]) & 0xFF
    if op == OP_SPHERE:
        return f[pc, 0:3] - f[pc, 3], f[pc, 0:3] + f[pc, 3]
    if op == OP_MSPHERE:
        c0, c1, t0, t1, r = f[pc, 0:3], f[pc + 1, 0:3], f[pc + 1, 3], f[pc + 1, 4], f[pc, 3]
        ca = c0 + ((F(ts) - t0) / (t1 - t0)) * (c1 - c0)
        cb = c0 + ((F(te) - t0) / (t1 - t0)) * (c1 - c0)
        return np.minimum(ca - r, cb - r), np.maximum(ca + r, cb + r)
    if op == OP_CUBOID:
        return f[pc, 0:3], f[pc, 4:7]
    a0, a1, b0, b1, k = f[pc, 0:5]
    lo, hi = k - F(0.0001), k + F(0.0001)
    if op == OP_RECT_XY:
        return np.array([a0, b0, lo], F), np.array([a1, b1, hi], F)
    if op == OP_RECT_YZ:
        return np.array([lo, a0, b0], F), np.array([hi, a1, b1], F)
    return np.array([a0, lo, b0], F), np.array([a1, hi, b1], F)


def walk_tree(ops, f, nodes, base, o, d, time, tmin, closest, best_pc, stats=None, ts=0.0, te=1.0):
    """OP_BVH tree walk for ONE ray (hrt_types.h Bvh2Node): returns (closest, best_pc, prim id or None).  An exact tie
    is settled as the reference settles it: the later leaf of the reference's order wins iff the reference would still
    have reached it, i.e. iff its own bounding box passes Aabb::hit with t_max = t (hrt_device.cuh tie_goes_to_later)."""
    boxes = nodes[:, [0, 1, 2, 3, 4, 5, 8, 9, 10, 11, 12, 13]].copy().view(np.float16).astype(F)  # (n, 12)
    links = nodes.view(np.int32)[:, [3, 7]]
    inv = F(1.0) / d
    prim = None
    stack = []
    ref = 0
    while True:
        if ref >= 0:
            n = base + ref
            if stats is not None:
                stats[0] += 2
            llo, lhi = _slab1(o, inv, boxes[n, 0:6], tmin, closest)
            rlo, rhi = _slab1(o, inv, boxes[n, 6:12], tmin, closest)
            hl, hr = not (lhi < llo), not (rhi < rlo)
            if hl and hr:
                right_first = rlo < llo
                stack.append((int(links[n, 0]) if right_first else int(links[n, 1]), llo if right_first else rlo))
                ref = int(links[n, 1]) if right_first else int(links[n, 0])
                continue
            if hl or hr:
                ref = int(links[n, 0]) if hl else int(links[n, 1])
                continue
        else:
            pc = ~ref
            hit, t, pid = _leaf1(ops, f, pc, o, d, time, tmin, closest)
            if hit:
                take = t < closest or best_pc < 0
                if not take and not (t > closest):  # exact tie
                    later = max(pc, best_pc)
                    mn, mx = _ref_leaf_box(ops, f, later, ts, te)
                    later_wins = bool(_box(o[None, :], d[None, :], mn, mx, np.asarray([[tmin]], F), np.asarray([t], F), True)[0])
                    take = (pc > best_pc) == later_wins
                if take:
                    closest, best_pc, prim = t, pc, pid
        while True:
            if not stack:
                return closest, best_pc, prim
            ref, t_entry = stack.pop()
            if not (t_entry > closest):
                break


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


def _apply_push(ops, f, q, o, d):
    """One OP_TRANSLATE / OP_ROTATE record applied to copies of (o, d) (translation.rs:25-29, rotation.rs:103-116)."""
    o, d = o.copy(), d.copy()
    if (int(ops[q, 7]) & 0xFF) == OP_TRANSLATE:
        o = o - f[q, 0:3]
    else:
        sn, cs, axis = f[q, 0], f[q, 1], int(ops[q, 2])
        ia, ib = (axis + 1) % 3, (axis + 2) % 3
        for v in (o, d):
            va, vb = v[ia].copy(), v[ib].copy()
            v[ia] = cs * va + sn * vb
            v[ib] = -sn * va + cs * vb
    return o, d


def trace_stream(ops, rays, nodes=None, stats=None, spans=None, xi=None, pc_begin=0, pc_end=None):
    """Closest hit of every ray against the op stream; returns (hit mask, t, prim id).  `nodes`: the tree-node table of
    the fast form; `stats[0]` counts box tests.  `spans` (hrt_scene_get_tree_spans) with the WAVE form of the stream: walk
    as the wavefront render's stream walk does — at an OP_BVH_PRE record the tree's closest hit over [tmin, +inf), computed
    on its own (the tree stage), is merged, and the walk goes on behind the records that exist only for the tree.
    `xi`: one uniform per ray for ConstantMedium::hit (constant_medium.rs:34-76, restated below); without it media are
    stepped over.  `[pc_begin, pc_end)`: the part of the stream to walk (a medium's boundary sub-stream)."""
    f = ops.view(np.float32)
    n = len(rays)
    o = rays["o"].astype(F).copy()
    d = rays["d"].astype(F).copy()
    time = rays["time"].astype(F)
    tmin = rays["tmin"].astype(F)
    closest = rays["tmax"].astype(F).copy()
    prim = np.full(n, -1, dtype=np.int64)
    best_pc = np.full(n, -1, dtype=np.int64)
    pc_of = np.full(n, pc_begin, dtype=np.int64)
    saved = []  # ray-space stack: (ray indices, their o, their d) per pushed level
    pc = pc_begin
    pc_stop = len(ops) if pc_end is None else pc_end
    with np.errstate(all="ignore"):
        while pc < pc_stop:
            idx = np.nonzero(pc_of == pc)[0]
            w7 = int(ops[pc, 7])
            op, payload = w7 & 0xFF, w7 >> 8
            nxt = pc + 1
            if op == OP_END:
                break
            if op == OP_BVH_PRE:  # the WAVE form: w0 = index of the tree among the pre-walked ones, payload = to_pc
                span = spans[int(ops[pc, 0])]
                assert int(span[2]) == pc and int(span[3]) == payload and int(ops[pc, 1]) == int(span[1])
                bvh_pc, to_pc = int(span[0]), payload
                assert f[pc, 4] == f[bvh_pc, 4] and f[pc, 5] == f[bvh_pc, 5]
                for i in idx:
                    oi, di = o[i], d[i]
                    for q in range(pc, bvh_pc):  # the ray spaces entered between from_pc and the tree
                        if (int(ops[q, 7]) & 0xFF) in (OP_TRANSLATE, OP_ROTATE):
                            oi, di = _apply_push(ops, f, q, oi, di)
                    ts, te = f[bvh_pc, 4], f[bvh_pc, 5]
                    tt, leaf, pid = walk_tree(ops, f, nodes, int(ops[bvh_pc, 0]), oi, di, time[i], tmin[i], F(np.inf), -1, stats, ts, te)
                    if leaf < 0:
                        continue
                    take = tt < closest[i]
                    if not take and not (tt > closest[i]):  # exact tie with an earlier record (hrt_device.cuh traverse_uniform)
                        take = best_pc[i] < 0
                        if not take:
                            mn, mx = _ref_leaf_box(ops, f, leaf, ts, te)
                            take = bool(_box(oi[None, :], di[None, :], mn, mx, np.asarray([[tmin[i]]], F), np.asarray([tt], F), True)[0])
                    if take:
                        closest[i], best_pc[i], prim[i] = tt, leaf, pid
                pc_of[idx] = to_pc
                pc += 1
                continue
            if idx.size:
                oo, dd, cl, tm = o[idx], d[idx], closest[idx], tmin[idx]
                if op in (OP_BOX, OP_BOX_LOOSE):
                    if stats is not None:
                        stats[0] += idx.size
                    hit = _box(oo, dd, f[pc, 0:3], f[pc, 4:7], tm[:, None] if op == OP_BOX_LOOSE else tm, cl, op == OP_BOX_LOOSE)
                    pc_of[idx] = np.where(hit, pc + 1, payload)
                    pc += 1
                    continue
                if op == OP_BVH:
                    for i in idx:
                        closest[i], best_pc[i], pid = walk_tree(ops, f, nodes, int(ops[pc, 0]), o[i], d[i], time[i], tmin[i],
                                                                closest[i], best_pc[i], stats, f[pc, 4], f[pc, 5])
                        if pid is not None:
                            prim[i] = pid
                    pc_of[idx] = payload
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
                elif op in (OP_MEDIUM, OP_MEDIUM_SPHERE, OP_MEDIUM_CUBOID):
                    if xi is not None:  # constant_medium.rs:34-76 on the boundary sub-stream [pc + 1, payload)
                        sub = rays[idx].copy()
                        sub["o"], sub["d"] = oo, dd  # the ray in the ray space the medium lives in
                        sub["tmin"], sub["tmax"] = -np.inf, np.inf
                        h1, t1, _ = trace_stream(ops, sub, nodes, stats, None, None, pc + 1, payload)
                        sub["tmin"] = t1 + F(0.0001)
                        h2, t2, _ = trace_stream(ops, sub, nodes, stats, None, None, pc + 1, payload)
                        r1 = np.where(t1 < tm, tm, t1)
                        r2 = np.where(t2 > cl, cl, t2)
                        ok = h1 & h2 & ~(r1 >= r2)
                        r1 = np.where(r1 < 0, F(0), r1)
                        length = np.sqrt((dd[:, 0] * dd[:, 0] + dd[:, 1] * dd[:, 1]) + dd[:, 2] * dd[:, 2])
                        inside = (r2 - r1) * length
                        hit_distance = f[pc, 0] * np.log(xi[idx].astype(F))  # w0 = -1 / density
                        ok &= ~(hit_distance > inside)
                        t = r1 + hit_distance / length
                        sel = idx[ok]
                        closest[sel] = t[ok]
                        prim[sel] = int(ops[pc, 3])
                        best_pc[sel] = pc
                    pc_of[idx] = payload  # go on behind the boundary sub-stream
                    pc += 1
                    continue
                else:
                    raise AssertionError(f"unknown opcode {op:#x} at {pc}")
                sel = idx[ok]
                closest[sel] = t[ok]
                prim[sel] = pid
                best_pc[sel] = pc
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


def _both_forms(pkg, orc, world):
    gb, ob, _, _ = build_both(pkg, orc, world)
    N = pkg.native
    return gb, ob, gb.ops(N.HRT_STREAM_REFERENCE), gb.ops(N.HRT_STREAM_FAST), gb.tree_nodes()


@pytest.mark.parametrize("name", ["random", "cornell", "final"])
def test_fast_form_finds_the_same_hits(pkg, orc, name):
    """The FAST form (include/hrt.h): sound BVHs of plain primitives become OP_BVH trees — surface-area-heuristic
    topology, nearer child first — and must return the oracle's hits INCLUDING the primitive named on an exact tie
    between coincident surfaces (`final`'s ground boxes share faces): the leaf records stay in the reference's order
    and an equal-t hit only replaces an earlier record's.  BVHs with an unsound (axis-swapped ZX rect) leaf box keep the
    reference form, clipping behaviour included (Cornell)."""
    spec = pkg.make_scene(name, seed=4)
    gb, ob, ref_ops, fast_ops, nodes = _both_forms(pkg, orc, spec.world)
    i = gb.info()
    kinds, ref_kinds = fast_ops[:, 7] & 0xFF, ref_ops[:, 7] & 0xFF
    if name == "cornell":
        # its one BVH holds the ZX light (an unsound box): no tree.  What the fast form leaves out are SOUND inner boxes
        # only (they spare a warp nothing, hrt_scene.cpp emit_bvh); every loose box and every leaf box stays in place
        assert i.n_bvh_trees == 0 and i.n_tree_nodes == 0
        assert (kinds == OP_BOX_LOOSE).sum() == (ref_kinds == OP_BOX_LOOSE).sum() == 2
        assert 0 < (ref_kinds == OP_BOX).sum() - (kinds == OP_BOX).sum() == len(ref_ops) - len(fast_ops)
        assert kinds[kinds != OP_BOX].tolist() == ref_kinds[ref_kinds != OP_BOX].tolist()
    else:
        assert i.n_bvh_trees == {"random": 1, "final": 2}[name]
        assert i.n_fast_ops == len(fast_ops) and i.n_tree_nodes == len(nodes) and 0 < i.max_tree_depth <= 48
        # the trees replace the box records of those BVHs: one root box + one OP_BVH record per tree instead of 2n-1 boxes
        assert (kinds == OP_BVH).sum() == i.n_bvh_trees and i.n_fast_box_ops < i.n_box_ops // 10
    for k in (OP_BOX_LOOSE, OP_SPHERE, OP_MSPHERE, OP_CUBOID, OP_RECT_XY, OP_RECT_YZ, OP_RECT_ZX, OP_MEDIUM, OP_MEDIUM_SPHERE, OP_MEDIUM_CUBOID):
        assert (kinds == k).sum() == (ref_kinds == k).sum()
    rays = _rays(orc, ob, spec, n=700)
    want = ob.trace_hits(rays, np.full(len(rays), 0.5, dtype=np.float32))
    hit, t, prim = trace_stream(fast_ops, rays, nodes)
    medium = (want["hit"] == 1) & np.all(want["n"] == 0.0, axis=1)
    keep = ~medium & np.isfinite(want["t"])
    assert np.array_equal(hit[keep], want["hit"][keep] == 1)
    m = keep & (want["hit"] == 1)
    assert m.sum() > len(rays) // 10
    assert np.array_equal(prim[m], want["prim_id"][m]), (name, int((prim[m] != want["prim_id"][m]).sum()))
    assert np.allclose(t[m], want["t"][m], rtol=1e-5, atol=0)


@pytest.mark.parametrize("name,form", [("cornell-smoke", "reference"), ("cornell-smoke", "fast"), ("final", "fast"), ("final", "wave")])
def test_medium_semantics_match_the_oracle(pkg, orc, name, form):
    """ConstantMedium::hit restated in the numpy interpreter (two boundary queries over the medium's sub-stream, the
    clamps, the exponential free path with an injected uniform) against the oracle with the same uniforms: every hit —
    surfaces AND media — names the same primitive at the same t, on each form of the stream.  (logf differs by an ulp
    between numpy and glibc: media get 1e-5.)"""
    N = pkg.native
    spec = pkg.make_scene(name, seed=4)
    gb, ob, ref_ops, fast_ops, nodes = _both_forms(pkg, orc, spec.world)
    ops = {"reference": ref_ops, "fast": fast_ops, "wave": gb.ops(N.HRT_STREAM_WAVE)}[form]
    rays = _rays(orc, ob, spec, n=400)
    rng = np.random.default_rng(12)
    xi = rng.uniform(0.02, 0.98, len(rays)).astype(np.float32)
    want = ob.trace_hits(rays, xi)
    hit, t, prim = trace_stream(ops, rays, nodes, spans=gb.tree_spans(), xi=xi)
    fin = np.isfinite(want["t"]) | (want["hit"] == 0)
    assert fin.sum() > len(rays) * 0.9
    assert np.array_equal(hit[fin], want["hit"][fin] == 1)
    m = fin & (want["hit"] == 1)
    medium = m & np.all(want["n"] == 0.0, axis=1)
    assert medium.sum() > 10, int(medium.sum())
    assert np.array_equal(prim[m], want["prim_id"][m]), (name, form, int((prim[m] != want["prim_id"][m]).sum()))
    assert np.allclose(t[m], want["t"][m], rtol=1e-5, atol=0)


@pytest.mark.parametrize("name", ["random", "final", "nested"])
def test_tree_spans_step_over_only_what_belongs_to_the_tree(pkg, orc, name):
    """hrt_scene_get_tree_spans (hrt_types.h PreTree::from_pc / to_pc): the records of a span besides the tree itself are
    sound boxes that cover nothing else and balanced ray-space pushes / pops, and a stream walk that merges each tree's own
    closest hit at from_pc and goes on at to_pc — what wave_trace_kernel does — finds the oracle's hits, primitive ids
    included."""
    world = nested_tree_world(pkg) if name == "nested" else pkg.make_scene(name, seed=4).world
    gb, ob, ref_ops, ops, nodes = _both_forms(pkg, orc, world)
    spans = gb.tree_spans()
    assert len(spans) == gb.info().n_bvh_trees >= 1
    kinds = ops[:, 7] & 0xFF
    for bvh_pc, ctx, a, b in spans.tolist():
        assert kinds[bvh_pc] == OP_BVH and kinds[bvh_pc - 1] == OP_BOX and 0 <= a < bvh_pc
        end = int(ops[bvh_pc, 7]) >> 8
        assert end <= b <= len(ops)
        front, back = kinds[a:bvh_pc - 1], kinds[end:b]
        assert all(k in (OP_BOX, OP_TRANSLATE, OP_ROTATE) for k in front) and all(k == OP_POP for k in back)
        assert sum(k in (OP_TRANSLATE, OP_ROTATE) for k in front) == len(back)
        for q in range(a, bvh_pc):
            if kinds[q] == OP_BOX:  # every box of the span closes at the span's end or at the tree's
                assert (int(ops[q, 7]) >> 8) in (b, end) or kinds[(int(ops[q, 7]) >> 8):b].tolist() == [OP_POP] * (b - (int(ops[q, 7]) >> 8))
        # nothing jumps INTO the span: no earlier box skips to a record inside it
        skips = np.array([int(w) >> 8 for w, k in zip(ops[:a, 7], kinds[:a]) if k in (OP_BOX, 0x11)])
        assert not np.any((skips > a) & (skips < b))
    if name == "final":  # the ground boxes: root box + tree; the sphere cluster: a leaf box and Translation(Rotation(.)) around them
        ends = np.array([int(ops[p, 7]) >> 8 for p in spans[:, 0]])
        assert (spans[:, 0] - 1 - spans[:, 2]).tolist() == [0, 3] and (spans[:, 3] - ends).tolist() == [0, 2]
    if name == "nested":
        rng = np.random.default_rng(3)
        n = 600
        rays = make_rays(orc, rng.uniform(-12, 12, (n, 3)).astype(np.float32) + np.array([0, 6, 0], np.float32),
                         rng.normal(size=(n, 3)).astype(np.float32) - np.array([0, 0.6, 0], np.float32), time=rng.random(n, dtype=np.float32))
    else:
        rays = _rays(orc, ob, pkg.make_scene(name, seed=4), n=500)
    want = ob.trace_hits(rays, np.full(len(rays), 0.5, dtype=np.float32))
    wave = gb.ops(pkg.native.HRT_STREAM_WAVE)
    pre = np.nonzero((wave[:, 7] & 0xFF) == OP_BVH_PRE)[0]
    assert pre.tolist() == sorted(spans[:2, 2].tolist()) and len(wave) == len(ops)  # the first kMaxPreTrees (2) trees
    assert np.array_equal(np.delete(wave, pre, axis=0), np.delete(ops, pre, axis=0)) and np.all((ops[pre, 7] & 0xFF) == OP_BOX)
    hit, t, prim = trace_stream(wave, rays, nodes, spans=spans)
    medium = (want["hit"] == 1) & np.all(want["n"] == 0.0, axis=1)
    keep = ~medium & np.isfinite(want["t"])
    assert np.array_equal(hit[keep], want["hit"][keep] == 1)
    m = keep & (want["hit"] == 1)
    assert m.sum() > len(rays) // 10
    assert np.array_equal(prim[m], want["prim_id"][m]), (name, int((prim[m] != want["prim_id"][m]).sum()))
    assert np.allclose(t[m], want["t"][m], rtol=1e-5, atol=0)


def test_fast_form_leaves_out_only_sound_inner_boxes_of_small_subtrees(pkg, orc):
    """hrt_scene.cpp emit_bvh: in a BVH that is not a tree (its leaves are translated cuboids) the fast form drops the sound
    inner boxes with at most 16 leaves beneath and keeps the larger ones, every leaf box and every loose box; the hits —
    primitive ids included — are the oracle's, and the library's own count of box records is what the stream holds."""
    S = pkg.scene
    rng = np.random.default_rng(17)
    white = S.Lambertian(S.SolidColor((0.7, 0.7, 0.7)))
    objs = []
    for _ in range(40):
        lo = rng.uniform(-1.0, 0.0, 3)
        objs.append(S.Translation(S.Cuboid(tuple(lo), tuple(lo + rng.uniform(0.3, 1.2, 3)), white), tuple(rng.uniform(-9, 9, 3))))
    objs.append(S.Rect(S.Plane.ZX, -3.0, -1.0, 1.0, 4.0, 9.5, S.DiffuseLight(S.SolidColor((4.0, 4.0, 4.0)))))  # an unsound leaf box
    world = S.BvhNode(objs, 0.0, 1.0)
    gb, ob, ref_ops, ops, nodes = _both_forms(pkg, orc, world)
    kinds, ref_kinds = ops[:, 7] & 0xFF, ref_ops[:, 7] & 0xFF
    i = gb.info()
    assert i.n_bvh_trees == 0 and i.n_fast_box_ops == int(((kinds == OP_BOX) | (kinds == OP_BOX_LOOSE)).sum())
    n_leaves = len(objs)
    assert (ref_kinds == OP_BOX).sum() + (ref_kinds == OP_BOX_LOOSE).sum() == 2 * n_leaves - 1
    assert (kinds == OP_BOX_LOOSE).sum() == (ref_kinds == OP_BOX_LOOSE).sum() >= 1
    assert kinds[kinds != OP_BOX].tolist() == ref_kinds[ref_kinds != OP_BOX].tolist()  # nothing else moved
    # every leaf box is still there: a box record right in front of each leaf's first record
    first_of_leaf = [pc for pc in range(len(ops)) if kinds[pc] in (OP_TRANSLATE, OP_RECT_ZX) and (pc == 0 or kinds[pc - 1] != OP_TRANSLATE)]
    assert len(first_of_leaf) == n_leaves and all(kinds[pc - 1] in (OP_BOX, OP_BOX_LOOSE) for pc in first_of_leaf)
    # the inner boxes that stay cover more than 16 leaves (or are loose); some were dropped, some kept
    leaf_pcs = np.array(first_of_leaf)
    inner = [pc for pc in range(len(ops)) if kinds[pc] == OP_BOX and pc + 1 not in first_of_leaf]
    assert 0 < len(inner) < n_leaves - 1 - int((kinds == OP_BOX_LOOSE).sum())
    for pc in inner:
        assert ((leaf_pcs > pc) & (leaf_pcs < (int(ops[pc, 7]) >> 8))).sum() > 16
    org = rng.uniform(-14, 14, (1500, 3)).astype(np.float32)
    rays = make_rays(orc, org, rng.uniform(-9, 9, (1500, 3)).astype(np.float32) - org, time=rng.random(1500, dtype=np.float32))
    want = ob.trace_hits(rays, np.full(len(rays), 0.5, dtype=np.float32))
    for form in (ref_ops, ops):
        hit, t, prim = trace_stream(form, rays, nodes)
        assert np.array_equal(hit, want["hit"] == 1)
        m = want["hit"] == 1
        assert m.sum() > 80 and np.array_equal(prim[m], want["prim_id"][m]) and np.allclose(t[m], want["t"][m], rtol=1e-5, atol=0)


def test_medium_records_name_their_boundary_shape(pkg, orc):
    """A ConstantMedium whose boundary is one plain sphere / one cuboid (bare or inside one run of ray-space pushes) is
    flagged for the closed-form paths (hrt_types.h OP_MEDIUM_SPHERE / OP_MEDIUM_CUBOID); its sub-stream stays in place."""
    kinds = lambda name: [int(x) & 0xFF for x in _both_forms(pkg, orc, pkg.make_scene(name, 1).world)[3][:, 7]]  # noqa: E731
    smoke = kinds("cornell-smoke")
    assert smoke.count(OP_MEDIUM_CUBOID) == 2 and smoke.count(OP_MEDIUM) == 0
    i = smoke.index(OP_MEDIUM_CUBOID)
    assert smoke[i + 1:i + 6] == [OP_TRANSLATE, OP_ROTATE, OP_CUBOID, OP_POP, OP_POP]
    final = kinds("final")
    assert final.count(OP_MEDIUM_SPHERE) == 2 and final.count(OP_MEDIUM) == 0 and final.count(OP_MEDIUM_CUBOID) == 0
    S = pkg.scene
    m = S.Lambertian(S.SolidColor((0.5, 0.5, 0.5)))
    bare = S.ConstantMedium(S.Cuboid((0, 0, 0), (1, 1, 1), m), 0.5, S.SolidColor((1, 1, 1)))
    lst = S.ConstantMedium(S.List([S.Cuboid((0, 0, 0), (1, 1, 1), m), S.Sphere((3, 0, 0), 1.0, m)]), 0.5, S.SolidColor((1, 1, 1)))
    ks = [int(x) & 0xFF for x in _both_forms(pkg, orc, S.List([bare, lst]))[3][:, 7]]
    assert ks.count(OP_MEDIUM_CUBOID) == 1 and ks.count(OP_MEDIUM) == 1  # a two-object boundary stays generic


def test_tree_structure_is_well_formed(pkg, orc):
    """Every leaf record of a tree is referenced exactly once, links stay inside the tree, the leaf records sit between the
    OP_BVH record and its end, and every child box (fp16, rounded outward) contains the boxes beneath it."""
    spec = pkg.make_scene("final", seed=4)
    gb, ob, ref_ops, ops, nodes = _both_forms(pkg, orc, spec.world)
    f = ops.view(np.float32)
    boxes = nodes[:, [0, 1, 2, 3, 4, 5, 8, 9, 10, 11, 12, 13]].copy().view(np.float16).astype(F)
    links = nodes.view(np.int32)[:, [3, 7]]
    for pc in np.where((ops[:, 7] & 0xFF) == OP_BVH)[0]:
        base, n_nodes, n_leaves, depth = (int(x) for x in ops[pc, 0:4].view(np.int32))
        end = int(ops[pc, 7]) >> 8
        assert (int(ops[pc - 1, 7]) & 0xFF) == OP_BOX and (int(ops[pc - 1, 7]) >> 8) == end  # root box skips the tree
        assert n_nodes == n_leaves - 1
        seen_leaf, seen_node = [], []

        def extent(pcl):
            op = int(ops[pcl, 7]) & 0xFF
            if op == OP_SPHERE:
                return np.concatenate([f[pcl, 0:3] - f[pcl, 3], f[pcl, 0:3] + f[pcl, 3]])
            assert op == OP_CUBOID
            return np.concatenate([f[pcl, 0:3], f[pcl, 4:7]])

        def visit(ref, d):
            if ref < 0:
                assert pc < ~ref < end
                seen_leaf.append(~ref)
                return extent(~ref), d
            assert 0 <= ref < n_nodes
            seen_node.append(ref)
            n = base + ref
            out, dmax = None, d
            for side in (0, 1):
                e, dd = visit(int(links[n, side]), d + 1)
                b = boxes[n, 6 * side:6 * side + 6]
                assert np.all(b[0:3] <= e[0:3]) and np.all(b[3:6] >= e[3:6])
                # ... and not by more than fp16 rounding
                assert np.all(np.abs(b - e) <= np.maximum(np.abs(e) * 2.0 ** -10, 2.0 ** -14))
                out = e if out is None else np.concatenate([np.minimum(out[0:3], e[0:3]), np.maximum(out[3:6], e[3:6])])
                dmax = max(dmax, dd)
            return out, dmax

        _, dmax = visit(0, 1)
        assert sorted(seen_node) == list(range(n_nodes)) and len(set(seen_leaf)) == len(seen_leaf) == n_leaves
        assert dmax - 1 == depth
        # leaf records = every primitive record between the OP_BVH record and its end
        prim_pcs = [q for q in range(pc + 1, end) if (int(ops[q, 7]) & 0xFF) in (OP_SPHERE, OP_MSPHERE, OP_CUBOID, OP_RECT_XY, OP_RECT_YZ, OP_RECT_ZX)]
        assert sorted(seen_leaf) == prim_pcs


def test_trees_visit_fewer_boxes(pkg, orc):
    """The point of the fast form: box tests per ray, reference form (tight tests) vs trees, by the interpreter."""
    spec = pkg.make_scene("random", seed=4)
    gb, ob, ref_ops, fast_ops, nodes = _both_forms(pkg, orc, spec.world)
    rays = _rays(orc, ob, spec, n=150)
    a, b = [0], [0]
    h1 = trace_stream(ref_ops, rays, None, a)
    h2 = trace_stream(fast_ops, rays, nodes, b)
    assert np.array_equal(h1[2], h2[2]) and np.array_equal(h1[1], h2[1])
    assert b[0] < 0.5 * a[0], (a, b)


def test_tree_edge_cases(pkg, orc):
    """Small and nested BVHs, moving spheres, coincident surfaces (the tie rule), an unknown builder id, the option after
    commit."""
    S, N = pkg.scene, pkg.native
    m = S.Lambertian(S.SolidColor((0.5, 0.5, 0.5)))
    m2 = S.Metal((0.5, 0.5, 0.5), 0.0)
    one = S.BvhNode([S.Sphere((0, 0, -5), 1.0, m)], 0.0, 1.0)
    three = S.BvhNode([S.Sphere((3, 0, -5), 1.0, m), S.Sphere((-3, 0, -5), 1.0, m), S.Sphere((-3, 3, -5), 1.0, m)], 0.0, 1.0)
    # six leaves, two pairs of COINCIDENT surfaces: identical spheres, and cuboids sharing the face x = 1
    tied = S.BvhNode([S.MovingSphere((0, 3, -5), (0, 4, -5), 0.0, 1.0, 0.5, m), S.Sphere((0, -3, -5), 0.5, m),
                      S.Sphere((0, -3, -5), 0.5, m2), S.Sphere((0, -6, -9), 0.5, m),
                      S.Cuboid((0, 0, -9), (1, 1, -8), m), S.Cuboid((1, 0, -9), (2, 1, -8), m2)], 0.0, 1.0)
    world = S.BvhNode([one, three, tied, S.Sphere((0, -1000, 0), 990.0, m)], 0.0, 1.0)
    gb, ob, ref_ops, fast_ops, nodes = _both_forms(pkg, orc, world)
    i = gb.info()
    assert i.n_bvh_trees == 1 and i.n_tree_nodes == 5  # `tied` only: the others have fewer than four children / BVH leaves
    bad = pkg.HrtBackend()
    with pytest.raises(pkg.HrtError) as ei:
        bad.set_bvh_builder(7)
    assert ei.value.code == -1
    with pytest.raises(pkg.HrtError):
        gb.set_bvh_builder(N.HRT_BVH_REFERENCE)  # immutable after commit, like every builder call
    rng = np.random.default_rng(1)
    d = rng.normal(size=(800, 3)).astype(np.float32)
    d[np.abs(d) < 1e-3] = 1e-3
    rays = make_rays(orc, np.tile(np.float32([0, 0, 4]), (800, 1)), d, time=rng.random(800, dtype=np.float32))
    # rays aimed at the coincident surfaces: the identical spheres, and points ON the shared cuboid face seen from inside
    aim = np.float32([[0, -3, -5]]) + 0.3 * rng.normal(size=(200, 3)).astype(np.float32) - np.float32([[0, 0, 4]])
    face = np.stack([np.full(100, 1.0), rng.uniform(0.05, 0.95, 100), rng.uniform(-8.95, -8.05, 100)], axis=1).astype(np.float32)
    inside = np.float32([[0.5, 0.5, -8.5]])
    rays = np.concatenate([rays, make_rays(orc, np.tile(np.float32([0, 0, 4]), (200, 1)), aim),
                           make_rays(orc, np.tile(inside, (100, 1)), face - inside)])
    want = ob.trace_hits(rays, np.full(len(rays), 0.5, dtype=np.float32))
    for ops, nd in ((ref_ops, None), (fast_ops, nodes)):
        hit, t, prim = trace_stream(ops, rays, nd)
        assert np.array_equal(hit, want["hit"] == 1) and (want["hit"] == 1).sum() > 100
        k = want["hit"] == 1
        assert np.array_equal(prim[k], want["prim_id"][k]) and np.allclose(t[k], want["t"][k], rtol=1e-5, atol=0)
    # the tie cases were really exercised: both identical spheres' later twin and the later cuboid win
    ids = set(want["prim_id"][want["hit"] == 1].tolist())
    sph_ids = [int(x) for x in fast_ops[(fast_ops[:, 7] & 0xFF) == OP_SPHERE][:, 5]]
    assert len(ids & set(sph_ids)) >= 3


def test_builder_default_can_come_from_the_environment(pkg, monkeypatch):
    """HRT_BVH_BUILDER=reference (what `bench.py --bvh reference` sets) is the default of scenes created afterwards; an
    explicit hrt_scene_set_bvh_builder still wins."""
    N = pkg.native
    spec = pkg.make_scene("random", seed=2)
    monkeypatch.setenv("HRT_BVH_BUILDER", "reference")
    a = pkg.HrtBackend()
    pkg.scene.emit(spec.world, a)
    b = pkg.HrtBackend()
    b.set_bvh_builder(N.HRT_BVH_TREES)
    pkg.scene.emit(spec.world, b)
    monkeypatch.delenv("HRT_BVH_BUILDER")
    c = pkg.HrtBackend()
    pkg.scene.emit(spec.world, c)
    assert a.info().n_bvh_trees == 0 and b.info().n_bvh_trees == 1 and c.info().n_bvh_trees == 1
    assert np.array_equal(a.ops(N.HRT_STREAM_FAST), a.ops(N.HRT_STREAM_REFERENCE))
    assert np.array_equal(b.ops(N.HRT_STREAM_FAST), c.ops(N.HRT_STREAM_FAST))
    assert np.array_equal(a.ops(), c.ops())  # the reference form does not depend on the option
