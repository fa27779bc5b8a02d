"""Headless front end (hyper-ray-tracer_b200/__main__.py): same flags and defaults as the reference's clap parser
(src/arguments.rs:21-47); frame writers."""
import importlib
import os

import numpy as np
import pytest

from conftest import graft


def _cli(pkg):
    return importlib.import_module(graft.PKG_NAME + ".__main__")


def test_flags_and_defaults_mirror_arguments_rs(pkg):
    cli = _cli(pkg)
    a = cli.parse([])
    assert (a.width, a.height, a.samples, a.depth, a.tile_size, a.scene) == (1280, 720, 100, 10, 80, "random")
    a = cli.parse("--width 400 --height 225 --samples 7 --depth 50 --tile-size 40 --scene cornell-smoke".split())
    assert (a.width, a.height, a.samples, a.depth, a.tile_size, a.scene) == (400, 225, 7, 50, 40, "cornell-smoke")
    # the eight `Scene` variants in clap's kebab-case (arguments.rs:10-19)
    assert sorted(pkg.SCENES) == sorted(["random", "two-spheres", "two-perlin-spheres", "earth", "simple-light", "cornell",
                                         "cornell-smoke", "final"])
    with pytest.raises(SystemExit):
        cli.parse(["--scene", "nope"])


def test_frame_writers(pkg, tmp_path):
    cli = _cli(pkg)
    h, w = 5, 7
    frame = np.zeros((h, w, 4), dtype=np.float32)
    frame[..., 3] = 1.0
    frame[0, :, 0] = 1.0  # bottom row red (rows are bottom-up, as in the reference's tiles)
    p = str(tmp_path / "f.npy")
    cli.write_frame(p, frame)
    assert np.array_equal(np.load(p), frame)
    p = str(tmp_path / "f.png")
    cli.write_frame(p, frame)
    from PIL import Image
    img = np.asarray(Image.open(p))
    assert img.shape == (h, w, 3) and img[-1, 0, 0] == 255 and img[0, 0, 0] == 0  # PNG is top-down
    p = str(tmp_path / "f.pfm")
    cli.write_frame(p, frame)
    raw = open(p, "rb").read()
    assert raw.startswith(b"PF\n7 5\n-1.0\n") and len(raw) == len(b"PF\n7 5\n-1.0\n") + h * w * 12


@pytest.mark.gpu
def test_cli_renders_a_frame(pkg, tmp_path):
    cli = _cli(pkg)
    out = str(tmp_path / "c.npy")
    rc = cli.main(f"--scene cornell --width 64 --height 64 --samples 32 --depth 50 --out {out}".split())
    assert rc == 0
    f = np.load(out)
    assert f.shape == (64, 64, 4) and np.all(f[..., 3] == 1.0) and np.nanmax(f[..., :3]) > 0.5


@pytest.mark.gpu
def test_cli_stored_scene_renders_the_same_frame(pkg, tmp_path):
    """--save-scene / --load-scene: the stored instance renders the same frame as the generated one (same render seed)."""
    cli = _cli(pkg)
    scene, a, b = str(tmp_path / "s.hrts"), str(tmp_path / "a.npy"), str(tmp_path / "b.npy")
    common = "--width 48 --height 32 --samples 16 --depth 20 --render-seed 5"
    assert cli.main(f"--scene random --seed 9 {common} --save-scene {scene} --out {a}".split()) == 0
    assert cli.main(f"--load-scene {scene} {common} --out {b}".split()) == 0
    assert np.allclose(np.nan_to_num(np.load(a)), np.nan_to_num(np.load(b)), rtol=2e-4, atol=2e-4)
    # ... and the Python harness's .npz instance format still loads (scene_io.py)
    spec = pkg.make_scene("random", 9)
    npz, c = str(tmp_path / "s.npz"), str(tmp_path / "c.npy")
    pkg.scene_io.save_scene(spec, npz)
    assert cli.main(f"--load-scene {npz} {common} --out {c}".split()) == 0
    assert np.allclose(np.nan_to_num(np.load(a)), np.nan_to_num(np.load(c)), rtol=2e-4, atol=2e-4)
