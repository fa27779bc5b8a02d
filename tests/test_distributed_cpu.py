"""Host-side logic of the spp-sharded multi-GPU path, on CPU with world_size-2 gloo (no GPU needed): slice arithmetic,
and the invariant the NCCL path relies on — disjoint slices all-reduced (sum) equal the single-rank accumulator."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def test_sample_slices_partition_the_range(pkg):
    f = pkg.renderer.sample_slice
    for samples in (1, 7, 100, 4096, 10000):
        for world in (1, 2, 3, 4, 8):
            parts = [f(samples, world, r) for r in range(world)]
            assert parts[0][0] == 0
            assert sum(c for _, c in parts) == samples
            for (b0, c0), (b1, _) in zip(parts, parts[1:]):
                assert b0 + c0 == b1
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, samples, out_path):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as graft
    pkg = graft.load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    begin, count = pkg.renderer.sample_slice(samples, world, rank)
    # stand-in for hrt_render_accum_device: a deterministic per-(pixel, sample) contribution, summed over the slice
    h, w = 6, 5
    pix = np.arange(h * w, dtype=np.float64).reshape(h, w, 1)
    acc = np.zeros((h, w, 4), dtype=np.float64)
    for s in range(begin, begin + count):
        acc[..., :3] += np.sin(0.1 * pix + s) ** 2 * np.array([1.0, 0.5, 0.25])
        acc[..., 3] += 1.0
    t = torch.from_numpy(acc)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)  # the ONE collective of the path
    if rank == 0:
        np.save(out_path, t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_equals_single_rank(tmp_path):
    samples = 37
    out = str(tmp_path / "acc.npy")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, samples, out), nprocs=2, join=True)
    got = np.load(out)
    h, w = 6, 5
    pix = np.arange(h * w, dtype=np.float64).reshape(h, w, 1)
    want = np.zeros((h, w, 4))
    for s in range(samples):
        want[..., :3] += np.sin(0.1 * pix + s) ** 2 * np.array([1.0, 0.5, 0.25])
        want[..., 3] += 1.0
    assert np.allclose(got, want, rtol=1e-12)
    assert np.all(got[..., 3] == samples)
