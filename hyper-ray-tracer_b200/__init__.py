"""hyper-ray-tracer_b200 — B200-native (sm_100a CUDA) implementation of hyper-ray-tracer's path-tracing hot path.

Layout:
  csrc/        hand-written CUDA kernels + the C-ABI library libhrt.so (include/hrt.h)
  native.py    ctypes binding of the C ABI (fails loudly when the CUDA extension is missing)
  scene.py     constructor-mirroring scene API of the reference (Sphere, Rect, BvhNode, Lambertian, ...)
  scenes.py    the reference's eight scene generators, seeded
  scene_io.py  scene instances on disk (.npz): save_scene / load_scene
  renderer.py  `Application::render` drop-in: single-GPU and spp-sharded multi-GPU (torch.distributed/NCCL)
"""
from . import native, renderer, scene, scene_io, scenes  # noqa: F401
from .native import HrtBackend, HrtError  # noqa: F401
from .scene import *  # noqa: F401,F403
from .scene_io import load_scene, save_scene  # noqa: F401
from .scenes import CONFIGS, SCENES, make_scene  # noqa: F401
