"""The scene library behind the C ABI: hrt_make_scene — the reference's eight generators
(src/application.rs:497-935) with an explicit seed — and hrt_scene_save / hrt_scene_load (scene-instance files).

The Python generators of hyper-ray-tracer_b200/scenes.py are an independent restatement of the same reference code on
numpy's PCG64; the C++ ones must issue the SAME builder calls in the SAME order — compared here as the bit-identical
flattened streams (both forms), material / texture tables included through the streams' ids — for every scene and
several seeds.  That also pins hrt_rng.hpp's restatement of numpy's SeedSequence + PCG64 + bounded integers.
GPU: hit records of the library-built scenes against the committed fixtures (the oracle's records, test_hit_fixtures.py).
"""
import os

import numpy as np
import pytest

from conftest import ROOT

SCENES = ["random", "two-spheres", "two-perlin-spheres", "earth", "simple-light", "cornell", "cornell-smoke", "final"]


def _library_scene(pkg, name, seed):
    gb = pkg.HrtBackend()
    image = pkg.scenes.load_earthmap() if name in ("earth", "final") else None
    root, view = gb.make_scene(name, seed, image)
    return gb, root, view


def _same_scene(pkg, a, b):
    N = pkg.native
    ia, ib = a.info(), b.info()
    for f, _ in ia._fields_:
        assert getattr(ia, f) == getattr(ib, f), f
    for which in (N.HRT_STREAM_REFERENCE, N.HRT_STREAM_FAST):
        assert np.array_equal(a.ops(which), b.ops(which))
    assert np.array_equal(a.tree_nodes(), b.tree_nodes())
    assert a.count() == b.count()


@pytest.mark.parametrize("name", SCENES)
@pytest.mark.parametrize("seed", [1, 4, 2**40 + 7])
def test_library_generators_equal_the_python_restatement(pkg, name, seed):
    if name in ("two-spheres", "earth", "cornell", "cornell-smoke") and seed != 1:
        pytest.skip("no random draws in this scene")
    spec = pkg.make_scene(name, seed)
    py = pkg.HrtBackend()
    pkg.scene.emit(spec.world, py)
    lib, root, view = _library_scene(pkg, name, seed)
    lib.commit(root)
    _same_scene(pkg, lib, py)
    # camera and background (application.rs:132-211)
    cam = spec.camera
    assert tuple(view.look_from) == tuple(np.float32(cam.look_from)) and tuple(view.look_at) == tuple(np.float32(cam.look_at))
    assert view.vfov == np.float32(cam.fov) and view.aperture == np.float32(cam.aperture)
    assert (view.focus_dist, view.time0, view.time1) == (10.0, 0.0, 1.0)
    assert tuple(view.background) == tuple(np.float32(spec.background))


def test_unknown_scene_and_missing_image(pkg):
    gb = pkg.HrtBackend()
    with pytest.raises(pkg.HrtError) as ei:
        gb.make_scene("cornell-box", 1)
    assert ei.value.code == -1 and "unknown scene" in ei.value.message
    # image_texture.rs:37-39: empty data gives (1, 0, 1); the generator still builds the scene
    g2 = pkg.HrtBackend()
    root, _ = g2.make_scene("earth", 1, None)
    g2.commit(root)
    assert g2.info().n_images == 0 and g2.info().n_textures == 1


@pytest.mark.parametrize("name", ["final", "cornell-smoke", "random"])
def test_scene_instance_files_round_trip(pkg, tmp_path, name):
    """save -> load re-issues the builder calls: same ids, same boxes, same trees, same flattened streams; the file of the
    reloaded scene is byte-identical."""
    lib, root, view = _library_scene(pkg, name, 3)
    p1, p2 = str(tmp_path / "a.hrts"), str(tmp_path / "b.hrts")
    lib.save(root, view, p1)
    again, root2, view2 = pkg.HrtBackend.load(p1)
    assert root2 == root and bytes(view2) == bytes(view)
    again.save(root2, view2, p2)
    assert open(p1, "rb").read() == open(p2, "rb").read()
    lib.commit(root)
    again.commit(root2)
    _same_scene(pkg, lib, again)
    # damaged files are refused, not crashed on
    data = open(p1, "rb").read()
    for bad in (data[:100], b"NOTASCENE" + data[9:], data[:len(data) // 2]):
        open(p2, "wb").write(bad)
        with pytest.raises(pkg.HrtError):
            pkg.HrtBackend.load(p2)
    with pytest.raises(pkg.HrtError):
        pkg.HrtBackend.load(str(tmp_path / "missing.hrts"))


@pytest.mark.gpu
@pytest.mark.parametrize("name", SCENES)
def test_library_built_scenes_give_the_oracles_hit_records(pkg, name, tmp_path):
    """The library's own generator (and a reload of its scene file) against the committed oracle hit records."""
    N = pkg.native
    z = np.load(os.path.join(ROOT, "tests", "golden", f"hits_{name}.npz"))
    rays, xi, want = z["rays"], z["xi"], z["hits"]
    lib, root, view = _library_scene(pkg, name, int(z["scene_seed"]))
    path = str(tmp_path / "s.hrts")
    lib.save(root, view, path)
    lib.commit(root)
    again, root2, _ = pkg.HrtBackend.load(path)
    again.commit(root2)
    m = want["hit"] == 1
    surf = m & np.isfinite(want["t"]) & np.isfinite(want["p"]).all(axis=1) & ~np.all(want["n"] == 0.0, axis=1)
    for gb in (lib, again):
        got = gb.trace_hits(rays, xi, flags=N.HRT_FLAG_EXACT_MATH)
        assert np.array_equal(got["hit"], want["hit"])
        for f in ("prim_id", "material_id", "face", "front_face"):
            assert np.array_equal(got[f][m], want[f][m]), f
        for f in ("t", "p", "n"):
            assert np.array_equal(got[f][surf], want[f][surf]), f
