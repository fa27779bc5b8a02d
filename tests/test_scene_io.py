"""Scene instances on disk (hyper-ray-tracer_b200/scene_io.py): a stored instance re-emits the identical
builder-call sequence, so the flattened op stream is bit-identical.  CPU only (builder + flattener run on the host)."""
import numpy as np
import pytest


def _stream(pkg, spec):
    b = pkg.HrtBackend()
    e = pkg.scene.emit(spec.world, b)
    i = b.info()
    return b.ops(), b.ops(1), b.tree_nodes(), (i.n_ops, i.n_materials, i.n_textures, i.n_noise_tables, i.n_images, i.n_media, i.n_contexts), e.root


@pytest.mark.parametrize("name", ["random", "two-spheres", "two-perlin-spheres", "earth", "simple-light", "cornell",
                                  "cornell-smoke", "final"])
def test_round_trip_is_bit_identical(pkg, tmp_path, name):
    spec = pkg.make_scene(name, seed=7)
    path = str(tmp_path / f"{name}.npz")
    pkg.save_scene(spec, path)
    back = pkg.load_scene(path)
    assert back.name == spec.name
    assert tuple(back.background) == tuple(float(x) for x in spec.background)
    assert back.camera == spec.camera or all(
        np.allclose(getattr(back.camera, f), getattr(spec.camera, f), rtol=0, atol=0) for f in
        ("look_from", "look_at", "fov", "aperture", "focus_dist", "time_0", "time_1"))
    ops0, fast0, nodes0, info0, root0 = _stream(pkg, spec)
    ops1, fast1, nodes1, info1, root1 = _stream(pkg, back)
    assert info0 == info1 and root0 == root1  # same sharing of materials / textures, same id allocation
    assert np.array_equal(ops0, ops1) and np.array_equal(fast0, fast1) and np.array_equal(nodes0, nodes1)


def test_second_generation_file_equals_first(pkg, tmp_path):
    spec = pkg.make_scene("final", seed=2)
    p1, p2 = str(tmp_path / "a.npz"), str(tmp_path / "b.npz")
    pkg.save_scene(spec, p1)
    pkg.save_scene(pkg.load_scene(p1), p2)
    with np.load(p1) as a, np.load(p2) as b:
        assert sorted(a.files) == sorted(b.files)
        for k in a.files:
            assert np.array_equal(a[k], b[k]), k


def test_foreign_and_future_files_are_rejected(pkg, tmp_path):
    import json
    p = str(tmp_path / "x.npz")
    np.savez(p, something=np.zeros(3))
    with pytest.raises(ValueError, match="not a hrt-scene file"):
        pkg.load_scene(p)
    doc = {"format": "hrt-scene", "version": 99}
    with open(p, "wb") as f:
        np.savez(f, scene_json=np.frombuffer(json.dumps(doc).encode(), dtype=np.uint8))
    with pytest.raises(ValueError, match="version 99"):
        pkg.load_scene(p)
