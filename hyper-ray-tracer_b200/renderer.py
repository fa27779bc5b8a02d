"""`Application::render` drop-in (/root/reference/src/application.rs:393-475) on B200.

Single GPU: one `hrt_render` call (path-trace kernel -> gamma resolve -> D2H of the RGBA-f32 frame the reference
uploads with glTexSubImage2D, application.rs:284-306).

Multi GPU (one process per GPU, torch.distributed/NCCL): samples-per-pixel shard across ranks — rank g renders the
disjoint Philox sample slice [begin_g, begin_g + count_g) of every pixel into its own fp32 accumulator, ONE
all-reduce(sum) over NVLink combines them, rank 0 resolves.  The union of the slices is exactly the 1-GPU sample
set, so the N-GPU image equals the 1-GPU image up to fp32 summation order.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from . import native
from .scene import SceneSpec, emit


def sample_slice(samples: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[begin, count) of rank's sample slice; the remainder is spread over the first ranks."""
    q, r = divmod(int(samples), int(world_size))
    begin = rank * q + min(rank, r)
    count = q + (1 if rank < r else 0)
    return begin, count


class Renderer:
    """Owns one committed scene on libhrt and renders it."""

    def __init__(self, spec: SceneSpec, device: int = 0, upload: bool = True, bvh_builder: Optional[int] = None):
        """bvh_builder: native.HRT_BVH_TREES / HRT_BVH_REFERENCE (include/hrt.h hrt_scene_set_bvh_builder); None keeps the
        library default (the reference's trees)."""
        self.spec = spec
        self.device = int(device)
        self.backend = native.HrtBackend()
        if bvh_builder is not None:
            self.backend.set_bvh_builder(bvh_builder)
        self.emitter = emit(spec.world, self.backend)
        if upload:
            self.backend.upload(self.device)

    # ---- single GPU, host buffers (the call a front end makes) ----
    def render(self, width: int, height: int, samples: int, depth: int, seed: int = 0, flags: int = 0,
               out: Optional[np.ndarray] = None, resolve: bool = True):
        return self.backend.render(self.spec.camera, width, height, samples, depth, self.spec.background, seed=seed,
                                   device=self.device, flags=flags, resolve=resolve, out=out)

    def render_progressive(self, width: int, height: int, samples: int, depth: int, batch: int, seed: int = 0, flags: int = 0,
                           on_frame=None):
        """Progressive delivery through hrt_render_progressive: the reference shows tiles as they finish
        (application.rs:284-306) and abandons the frame on a resize (:357-391); here the frame of the samples so far is
        delivered after every `batch` samples and `on_frame(done, total, frame)` may cancel by returning True.  Returns the
        list of (samples_done, frame copy) delivered and whether the render was cancelled.  The batches are disjoint Philox
        sample slices of ONE render, so the last frame equals `render(...)` up to f32 summation order."""
        frames = []

        def cb(done, total, frame):
            frames.append((done, frame.copy()))
            return bool(on_frame(done, total, frame)) if on_frame is not None else False
        _, _, cancelled = self.backend.render_progressive(self.spec.camera, width, height, samples, depth, self.spec.background,
                                                          batch, cb, seed=seed, device=self.device, flags=flags)
        return frames, cancelled

    # ---- device-resident pieces for the distributed path ----
    def render_slice_into(self, accum, width, height, samples, depth, seed, sample_begin, sample_count, stream_ptr=0,
                          flags=0, want_stats=False):
        """Adds the slice's radiance sums into `accum` (torch CUDA tensor [h, w, 4] f32 on self.device)."""
        return self.backend.render_accum_device(self.spec.camera, width, height, samples, depth, self.spec.background, seed,
                                                self.device, accum.data_ptr(), stream_ptr, sample_begin, sample_count, flags,
                                                want_stats)

    def resolve_into(self, accum, out, width, height, samples, stream_ptr=0):
        self.backend.resolve_device(self.device, accum.data_ptr(), width, height, samples, out.data_ptr(), stream_ptr)


class DistributedRenderer:
    """spp-sharded render over torch.distributed (backend nccl on GPUs).  Also runs with world_size 1."""

    def __init__(self, spec: SceneSpec, width: int, height: int, device: int, rank: int = 0, world_size: int = 1):
        import torch

        self.torch = torch
        self.rank, self.world_size = int(rank), int(world_size)
        self.width, self.height = int(width), int(height)
        self.r = Renderer(spec, device=device)
        dev = torch.device("cuda", device)
        self.accum = torch.zeros((height, width, 4), dtype=torch.float32, device=dev)
        self.rgba = torch.zeros((height, width, 4), dtype=torch.float32, device=dev)
        self.host = torch.empty((height, width, 4), dtype=torch.float32, pin_memory=True)

    def step(self, samples: int, depth: int, seed: int, to_host: bool = False, flags: int = 0, want_stats: bool = False):
        """One full frame: zero accum, render this rank's slice, all-reduce(sum), resolve (+ optional D2H on rank 0).
        Everything is enqueued on torch's current stream."""
        torch = self.torch
        stream = torch.cuda.current_stream().cuda_stream
        begin, count = sample_slice(samples, self.world_size, self.rank)
        self.accum.zero_()
        st = None
        if count > 0:
            st = self.r.render_slice_into(self.accum, self.width, self.height, samples, depth, seed, begin, count, stream,
                                          flags, want_stats)
        if self.world_size > 1:
            torch.distributed.all_reduce(self.accum, op=torch.distributed.ReduceOp.SUM)
        if self.rank == 0:
            self.r.resolve_into(self.accum, self.rgba, self.width, self.height, samples, stream)
            if to_host:
                self.host.copy_(self.rgba, non_blocking=True)
        return st
