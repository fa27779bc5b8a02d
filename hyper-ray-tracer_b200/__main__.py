"""Headless front end mirroring the reference's command line (/root/reference/src/arguments.rs:21-47):

    python -m hyper_ray_tracer_b200 --width 1280 --height 720 --samples 100 --depth 10 --scene random --out frame.png

(run it as `python hrt_cli.py ...` from the repo root: the package directory name contains a hyphen).  Same flags, same
defaults (`--tile-size` is accepted for compatibility; tiling is the GPU kernel's business).  The reference only shows
its result in a GLFW window and never writes a file; here the frame is written as PNG (8-bit, gamma
already applied by the resolve, top row first), `.npy` (float32 RGBA, bottom-up rows exactly as the reference's tiles)
or `.pfm`.  The "Rendered image in mm:ss" log line of application.rs:260-277 is reproduced.
"""
from __future__ import annotations

import argparse
import sys
import time

import numpy as np

from . import native, scene_io, scenes
from . import scene as scene_mod


def parse(argv=None):
    ap = argparse.ArgumentParser(prog="hyper-ray-tracer-b200", description="B200 path-tracing core behind hyper-ray-tracer's CLI")
    ap.add_argument("--width", type=int, default=1280, help="Width of the window")        # arguments.rs:25
    ap.add_argument("--height", type=int, default=720, help="Height of the window")      # arguments.rs:29
    ap.add_argument("--samples", type=int, default=100, help="Samples per pixel")         # arguments.rs:33
    ap.add_argument("--depth", type=int, default=10, help="Max depth")                    # arguments.rs:37
    ap.add_argument("--tile-size", type=int, default=80, help="Tile size (ignored)")      # arguments.rs:41
    ap.add_argument("--scene", default="random", choices=sorted(scenes.SCENES), help="Scene")  # arguments.rs:45
    ap.add_argument("--seed", type=int, default=1, help="scene-instance seed (the reference uses an unseeded thread_rng)")
    ap.add_argument("--render-seed", type=int, default=0, help="Philox key of the sample streams")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--out", default="frame.png", help="output file: .png, .npy (float32 RGBA, bottom-up) or .pfm")
    ap.add_argument("--bvh", default="trees", choices=["trees", "reference"],
                    help="flattened form that renders: sound BVHs as stack-walked surface-area-heuristic trees (default; same "
                         "hits, the reference's tie rule kept) or every BvhNode as the reference built it")
    ap.add_argument("--save-scene", default=None, metavar="FILE", help="write the generated scene instance (.npz) before rendering")
    ap.add_argument("--load-scene", default=None, metavar="FILE", help="render a stored scene instance instead of --scene/--seed")
    return ap.parse_args(argv)


def write_frame(path: str, rgba: np.ndarray) -> None:
    """rgba: (h, w, 4) float32, rows bottom-up, gamma-resolved (what hrt_render returns)."""
    if path.endswith(".npy"):
        np.save(path, rgba)
    elif path.endswith(".pfm"):
        rgb = np.ascontiguousarray(rgba[..., :3].astype("<f4"))  # PFM stores rows bottom-up already
        with open(path, "wb") as f:
            f.write(f"PF\n{rgb.shape[1]} {rgb.shape[0]}\n-1.0\n".encode())
            f.write(rgb.tobytes())
    else:
        from PIL import Image
        img = np.nan_to_num(rgba[::-1, :, :3], nan=0.0, posinf=1.0, neginf=0.0)
        Image.fromarray((np.clip(img, 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8), "RGB").save(path)


def main(argv=None) -> int:
    a = parse(argv)
    if native.device_count() < 1:
        print("error: no CUDA device — this renderer has no CPU fallback", file=sys.stderr)
        return 2
    print("Generating world...")                                    # application.rs:131
    builder = {"reference": native.HRT_BVH_REFERENCE, "trees": native.HRT_BVH_TREES}[a.bvh]
    if a.load_scene and a.load_scene.endswith(".npz"):  # the Python harness's instance format (scene_io.py)
        spec = scene_io.load_scene(a.load_scene)
        gb = native.HrtBackend()
        gb.set_bvh_builder(builder)
        scene_mod.emit(spec.world, gb)
        cam, background = spec.camera, spec.background
        if a.save_scene:
            scene_io.save_scene(spec, a.save_scene)
    else:
        # the library's own scene generators and scene-instance files (hrt_make_scene / hrt_scene_load, include/hrt.h)
        if a.load_scene:
            gb, root, view = native.HrtBackend.load(a.load_scene)
        else:
            gb = native.HrtBackend()
            root, view = gb.make_scene(a.scene, a.seed, scenes.load_earthmap() if a.scene in ("earth", "final") else None)
        if a.save_scene:
            gb.save(root, view, a.save_scene)
        gb.set_bvh_builder(builder)
        gb.commit(root)
        cam = scene_mod.Camera(tuple(view.look_from), tuple(view.look_at), view.vfov, view.aperture, view.focus_dist, view.time0,
                               view.time1)
        background = tuple(view.background)
    gb.upload(a.device)
    print("Generated world")                                        # application.rs:199
    print("Rendering image...")                                     # application.rs:387
    t0 = time.time()
    frame, st = gb.render(cam, a.width, a.height, a.samples, a.depth, background, seed=a.render_seed, device=a.device)
    dt = time.time() - t0
    print(f"Rendered image in {int(dt) // 60:02d}:{int(dt) % 60:02d}! ({dt * 1e3:.0f} ms, kernel {st.kernel_ms:.1f} ms)")  # :266-271
    print(f"  Width: {a.width}\n  Height: {a.height}\n  Samples: {a.samples}\n  Depth: {a.depth}\n  Objects: {gb.count()}")  # :272-277
    print(f"  Paths: {st.paths}  Rays: {st.rays}  ({st.paths / max(st.kernel_ms, 1e-6) / 1e3:.1f} Mpaths/s)")
    write_frame(a.out, frame)
    print(f"wrote {a.out}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
