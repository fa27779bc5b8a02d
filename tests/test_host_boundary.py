"""CPU-side tests of the drop-in boundary: libhrt.so loads, exports every symbol include/hrt.h declares, and its
host-side flattener reproduces the reference's construction-time arithmetic (checked against the oracle).
No compute calls here — those need a GPU (tests marked `gpu`)."""
import ctypes
import re
import os

import numpy as np
import pytest

from conftest import ROOT, build_both


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.native.load_library()
    header = open(os.path.join(ROOT, "include", "hrt.h")).read()
    declared = set(re.findall(r"\b(hrt_[a-z0-9_]+)\s*\(", header))
    declared -= {"hrt_scene"}  # type name
    assert declared, "no prototypes parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"libhrt.so does not export {name}"
    assert set(pkg.native.EXPORTS) == declared
    assert lib.hrt_abi_version() == pkg.native.ABI_VERSION == 3


def test_no_cpu_fallback(pkg):
    """Without a CUDA device every compute entry point must fail loudly (HRT_ERR_CUDA), never compute on the CPU."""
    if pkg.native.device_count() > 0:
        pytest.skip("a GPU is present")
    S = pkg.scene
    b = pkg.HrtBackend()
    S.emit(S.BvhNode([S.Sphere((0, 0, 0), 1.0, S.Dielectric(1.5))], 0.0, 1.0), b)
    cam = S.Camera((0, 0, 5), (0, 0, 0), 40.0, 0.0)
    with pytest.raises(pkg.HrtError) as ei:
        b.render(cam, 8, 8, 1, 5, (0, 0, 0))
    assert ei.value.code == -3 and "no CPU fallback" in ei.value.message
    rays = np.zeros(1, dtype=pkg.native.RAY_DTYPE)
    with pytest.raises(pkg.HrtError):
        b.trace_hits(rays)
    with pytest.raises(pkg.HrtError):
        b.upload(0)


def test_error_behaviour(pkg):
    S = pkg.scene
    b = pkg.HrtBackend()
    with pytest.raises(pkg.HrtError):
        b.mat_lambertian(5)  # unknown texture id
    with pytest.raises(pkg.HrtError):
        b.bvh([], 0.0, 1.0)  # reference panics "no elements in scene" (bvh_node.rs:38)
    with pytest.raises(pkg.HrtError):
        b.count()  # not committed
    t = b.tex_solid((1, 1, 1))
    m = b.mat_lambertian(t)
    s = b.sphere((0, 0, 0), 1.0, m)
    with pytest.raises(pkg.HrtError):
        b.rect(7, 0, 1, 0, 1, 0, m)
    with pytest.raises(pkg.HrtError):
        b.tree_spans()  # not committed
    b.commit(b.bvh([s], 0.0, 1.0))
    with pytest.raises(pkg.HrtError):
        b.sphere((0, 0, 0), 1.0, m)  # immutable after commit
    # a scene without trees: no spans, and its wave form is its fast form; an unknown form of the stream is an error
    N = pkg.native
    assert b.tree_spans().shape == (0, 4) and np.array_equal(b.ops(N.HRT_STREAM_WAVE), b.ops(N.HRT_STREAM_FAST))
    with pytest.raises(pkg.HrtError):
        b.ops(3)
    # ConstantMedium nested in a ConstantMedium boundary is rejected with a message
    b2 = pkg.HrtBackend()
    inner = S.ConstantMedium(S.Sphere((0, 0, 0), 1.0, S.Dielectric(1.5)), 0.1, S.SolidColor((1, 1, 1)))
    outer = S.ConstantMedium(inner, 0.1, S.SolidColor((1, 1, 1)))
    with pytest.raises(pkg.HrtError) as ei:
        S.emit(outer, b2)
    assert ei.value.code == -2


@pytest.mark.parametrize("name", ["random", "two-spheres", "two-perlin-spheres", "earth", "simple-light", "cornell",
                                  "cornell-smoke", "final"])
def test_flattener_matches_oracle_construction(pkg, orc, name):
    """BvhNode::new ordering, every bounding_box and world.count() agree with the oracle on all eight scenes."""
    spec = pkg.make_scene(name, seed=5)
    gb, ob, e1, e2 = build_both(pkg, orc, spec.world)
    assert gb.bvh_leaf_order(e1.root) == ob.bvh_leaf_order(e2.root)
    assert gb.count() == ob.count()
    for key, oid in e1.object_ids.items():
        try:
            want = ob.bounding_box(e2.object_ids[key])
        except orc.OracleError:
            continue
        assert np.array_equal(gb.bounding_box(oid), want), (name, oid)
    info = gb.info()
    assert info.n_box_ops == ob.bvh_node_count(e2.root) + sum(
        ob.bvh_node_count(e2.object_ids[k]) for k in e2.object_ids
        if e2.object_ids[k] != e2.root and ob.lib.orc_bvh_node_count(ob.handle, e2.object_ids[k]) > 0)


def test_stream_layout_cornell(pkg):
    """Op stream of the Cornell box: 15 boxes in DFS order, skip links forward, loose flags exactly on the light leaf
    and on its unsound ancestor {back, light} (Q2)."""
    N = pkg.native
    spec = pkg.make_scene("cornell", 1)
    gb = pkg.HrtBackend()
    e = pkg.scene.emit(spec.world, gb)
    ops = gb.ops()
    opc = ops[:, 7] & 0xFF
    payload = ops[:, 7] >> 8
    OP_END, OP_BOX, OP_BOX_LOOSE, OP_RECT_ZX, OP_CUBOID, OP_TRANSLATE, OP_ROTATE, OP_POP = 0x50, 0x10, 0x11, 0x32, 0x33, 0x40, 0x41, 0x42
    assert opc[-1] == OP_END
    boxes = np.where((opc == OP_BOX) | (opc == OP_BOX_LOOSE))[0]
    assert len(boxes) == 15
    for i in boxes:
        assert i < payload[i] <= len(ops) - 1  # skip links point forward, inside the stream
    loose = np.where(opc == OP_BOX_LOOSE)[0]
    assert len(loose) == 2 and gb.info().n_loose_boxes == 2
    light_id = e.object_ids[id(spec.world.objects[2])]
    light_pc = [i for i in np.where(opc == OP_RECT_ZX)[0] if ops[i, 6] == light_id]
    assert len(light_pc) == 1 and opc[light_pc[0] - 1] == OP_BOX_LOOSE  # the light's own leaf box
    # the other loose box is the parent of {back wall, light}: it spans z only from 227 (the swapped box) to 555
    parent = loose[0]
    mn = ops[parent, 0:3].view(np.float32)
    mx = ops[parent, 4:7].view(np.float32)
    assert mn[2] == 227.0 and mx[2] == np.float32(555.0) + np.float32(0.0001)
    # each rotated+translated box: TRANSLATE, ROTATE, CUBOID, POP, POP with context depth 2
    t = np.where(opc == OP_TRANSLATE)[0]
    assert len(t) == 2
    for i in t:
        assert list(opc[i:i + 5]) == [OP_TRANSLATE, OP_ROTATE, OP_CUBOID, OP_POP, OP_POP]
    assert gb.info().max_context_depth == 2 and gb.info().n_contexts == 5


def test_final_scene_stream(pkg):
    spec = pkg.make_scene("final", 1)
    gb = pkg.HrtBackend()
    pkg.scene.emit(spec.world, gb)
    i = gb.info()
    assert i.n_box_ops == 21 + 799 + 1999 and i.n_media == 2 and i.n_noise_tables == 1 and i.n_images == 1
    assert i.n_prim_ops == 400 + 1000 + 9
    assert i.n_loose_boxes >= 1  # the ZX light (123..423 x 147..412) has a swapped box
    assert i.time_min == 0.0 and i.time_max == 1.0


def test_tree_boxes_directed_rounding_extremes(pkg):
    """OP_BVH tree nodes carry their children's boxes in fp16 rounded OUTWARD (hrt_types.h Bvh2Node): by less than one
    fp16 step where fp16 can follow, to +-inf / the largest finite on the safe side beyond its range, to subnormals for
    tiny bounds.  (Structure and containment on the real scenes: tests/test_stream_interpreter.py.)"""
    S, N = pkg.scene, pkg.native
    gb = pkg.HrtBackend()
    m = S.Lambertian(S.SolidColor((0.5, 0.5, 0.5)))
    spheres = [((0.0, -100000.0, 0.0), 100000.0), ((1e-6, 1e-6, 1e-6), 3e-7), ((70000.0, 0.0, 0.0), 1.0), ((3.0, 4.0, 5.0), 0.7),
               ((-30.3, 1.1, 7.7), 2.2)]
    world = S.BvhNode([S.Sphere(c, r, m) for c, r in spheres], 0.0, 1.0)
    S.emit(world, gb)
    assert gb.info().n_bvh_trees == 1 and gb.info().n_tree_nodes == len(spheres) - 1
    ops, nodes = gb.ops(N.HRT_STREAM_FAST), gb.tree_nodes()
    f = ops.view(np.float32)
    boxes = nodes[:, [0, 1, 2, 3, 4, 5, 8, 9, 10, 11, 12, 13]].copy().view(np.float16)
    links = nodes.view(np.int32)[:, [3, 7]]
    seen = 0
    for n in range(len(nodes)):
        for side in (0, 1):
            ref = int(links[n, side])
            if ref >= 0:
                continue
            pc = ~ref
            c, r = f[pc, 0:3], f[pc, 3]
            mn, mx = c - r, c + r  # sphere.rs:77-83
            lo, hi = boxes[n, 6 * side:6 * side + 3], boxes[n, 6 * side + 3:6 * side + 6]
            assert np.all(lo.astype(np.float32) <= mn) and np.all(hi.astype(np.float32) >= mx)
            fin = np.isfinite(lo)
            with np.errstate(over="ignore"):
                assert np.all(np.nextafter(lo[fin], np.float16(np.inf)).astype(np.float32) > mn[fin])  # one step inward cuts in
            fin = np.isfinite(hi)
            assert np.all(np.nextafter(hi[fin], np.float16(-np.inf)).astype(np.float32) < mx[fin])
            seen += 1
    assert seen == len(spheres)
    h = boxes.astype(np.float32)
    assert np.isneginf(h).any() and np.isposinf(h).any()  # the r = 1e5 sphere
    assert (h == 65504.0).any()                            # min.x = 69999 of the far sphere: the largest finite fp16 is below it
    assert ((np.abs(h) < 1e-5) & (h != 0)).any()           # the tiny sphere: subnormals


def test_camera_init_matches_oracle(pkg, orc):
    gb = pkg.HrtBackend()
    ob = orc.OracleBackend()
    for name in ["random", "cornell", "final", "simple-light"]:
        cam = pkg.make_scene(name, 1).camera
        for (w, h) in [(400, 225), (600, 600), (801, 333)]:
            a = gb.camera_init(pkg.native.camera_desc(cam, w, h))
            b = ob.camera_init(orc.camera_desc(cam, w, h))
            assert bytes(a) == bytes(b)


def test_philox_kat(pkg):
    """Philox4x32-10 known answers (Random123 kat_vectors) through the library's own generator: with key = seed and
    counter = (pixel, sample, bounce<<8|block, 0x68727421) the uniforms are the top 24 bits of each word."""
    u = pkg.native.philox_uniforms(0x123456789ABCDEF0, 3, 7, 2, 1)
    assert u.shape == (4,) and np.all(u >= 0) and np.all(u < 1)

    def philox(ctr, key):
        M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
        c = list(ctr)
        k = list(key)
        for _ in range(10):
            p0 = M0 * c[0]
            p1 = M1 * c[2]
            c = [((p1 >> 32) ^ c[1] ^ k[0]) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c[3] ^ k[1]) & 0xFFFFFFFF,
                 p0 & 0xFFFFFFFF]
            k = [(k[0] + W0) & 0xFFFFFFFF, (k[1] + W1) & 0xFFFFFFFF]
        return c

    # Random123 known-answer vectors for philox4x32-10
    assert philox([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    want = philox([3, 7, (2 << 8) | 1, 0x68727421], [0x9ABCDEF0, 0x12345678])
    assert np.array_equal(u, np.array([(w >> 8) / 16777216.0 for w in want], dtype=np.float32))
