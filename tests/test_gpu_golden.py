"""Image parity against the committed golden frames (tests/golden/*.npz): the CPU oracle's own 4096-spp, depth-50 render
of every BASELINE config at half resolution (tools/make_goldens.py).  The GPU renders the same scene instance at many
more samples, so what remains is the GOLDEN's noise:
  * MAE <= 1/255 and PSNR >= 40 dB wherever the oracle's own noise floor allows it (C1, C2a, C2b);
  * for the low-light scenes (Cornell, Cornell-smoke, final) two independent 2048-spp oracle halves are only
    ~28 dB apart, so 4096 spp of the reference is itself ~6 dB away from converged; there the bar is "closer to the golden
    than the golden's two halves are to each other, by the 6 dB that the sample counts predict", plus unbiased z-scores.
Results are written to gpurun_out/golden_report.json (copied into profiles/ for the record)."""
import json
import os

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

GPU_SPP = {"C1": 32768, "C2a": 16384, "C2b": 16384, "C3": 16384, "C4": 16384, "C5": 16384}


def _load(cfg):
    path = os.path.join(ROOT, "tests", "golden", f"{cfg}.npz")
    if not os.path.exists(path):
        pytest.skip(f"golden {cfg} not generated")
    z = np.load(path)
    if "lin" in z.files:  # float32 LINEAR mean + its standard error (tools/make_goldens.py)
        lin = z["lin"].astype(np.float64)
        return np.sqrt(np.maximum(lin, 0.0)), lin, z["sigma"].astype(np.float64), json.loads(str(z["meta"])), True
    img = z["img"].astype(np.float64)  # first-version files: gamma-space float16
    return img, img ** 2, z["sigma"].astype(np.float64), json.loads(str(z["meta"])), False


@pytest.mark.parametrize("cfg", ["C1", "C2a", "C2b", "C3", "C4", "C5"])
def test_converged_image_matches_golden(pkg, cfg):
    img, gold_lin, sigma, meta, f32_golden = _load(cfg)
    spec = pkg.make_scene(meta["scene"], meta["scene_seed"])
    r = pkg.renderer.Renderer(spec, device=0)
    w, h, spp = meta["width"], meta["height"], GPU_SPP[cfg]
    acc, st = r.render(w, h, spp, meta["depth"], seed=4242, resolve=False)
    mean = np.nan_to_num(acc[..., :3].astype(np.float64) / spp, nan=0.0, posinf=0.0, neginf=0.0)
    gpu = np.sqrt(np.maximum(mean, 0.0))  # the reference's gamma resolve (application.rs:451-453)
    mae = float(np.abs(gpu - img).mean())
    mse = float(((np.clip(gpu, 0, 1) - np.clip(img, 0, 1)) ** 2).mean())
    psnr = float(10 * np.log10(1.0 / max(mse, 1e-12)))
    # pooled z-scores in linear space against the golden's own standard error
    k = 8
    hh, ww = h // k * k, w // k * k
    pool = lambda a: a[:hh, :ww].reshape(hh // k, k, ww // k, k, 3).sum(axis=(1, 3))  # noqa: E731
    var_blk = pool(sigma ** 2) * (1.0 + meta["spp"] / spp)
    ok = var_blk > 0
    z = (pool(mean) - pool(gold_lin))[ok] / np.sqrt(var_blk[ok])
    report = {"config": cfg, "scene": meta["scene"], "width": w, "height": h, "golden_spp": meta["spp"], "gpu_spp": spp,
              "mae": mae, "psnr_db": psnr, "oracle_half_vs_half_mae": meta["mae_half_vs_half"],
              "oracle_half_vs_half_psnr_db": meta["psnr_half_vs_half"], "z_rms": float(np.sqrt((z ** 2).mean())),
              "z_mean": float(z.mean()), "rays_per_path_gpu": st.rays / st.paths, "rays_per_path_oracle": meta["rays_per_path"],
              "gpu_mpaths_s": st.paths / (st.kernel_ms * 1e-3) / 1e6}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "golden_report.json")
    allr = json.load(open(path)) if os.path.exists(path) else {}
    allr[cfg] = report
    json.dump(allr, open(path, "w"), indent=1)
    print(json.dumps(report))
    # the two 2048-spp halves differ by noise of variance 4 sigma^2_4096; GPU-vs-golden carries ~sigma^2_4096
    # variance of (GPU - golden) = sigma^2_4096 * (1 + 4096 / gpu_spp); of (half A - half B) = 4 sigma^2_4096
    gain_db = 10.0 * np.log10(4.0 / (1.0 + meta["spp"] / spp))
    assert mae <= max(1.0 / 255.0, 1.08 * meta["mae_half_vs_half"] * 10 ** (-gain_db / 20.0)), report
    assert psnr >= min(40.0, meta["psnr_half_vs_half"] + gain_db - 0.6), report
    # Pooled z-scores against the golden's own standard error: unbiased (|mean| small) and of unit scale.  The float16
    # gamma-space files of the first generator version (C3 / C4 / C5: black background, every pixel noisy) pass as they
    # are; on C1 / C2a / C2b that storage put a 5.6e-4 relative error on the constant (0.7, 0.8, 1.0) background — many
    # sigmas of a nearly noise-free block, the +2 .. +22 z-means of round 1 — hence their float32 linear goldens.
    assert abs(report["z_mean"]) < 0.5, report
    assert report["z_rms"] < 1.6, report
    assert abs(report["rays_per_path_gpu"] - report["rays_per_path_oracle"]) < 0.01 * report["rays_per_path_oracle"], report


def test_two_gpu_nccl_render_equals_single_gpu(pkg):
    """spp sharding over NCCL (one process per GPU, torchrun): needs >= 2 GPUs, otherwise skipped."""
    import subprocess
    import sys
    if pkg.native.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29613", os.path.join(ROOT, "tools", "dist_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    assert json.loads(line)["ok"]


@pytest.mark.parametrize("cfg", ["C3", "C4", "C5"])
def test_deep_convergence_beyond_40db(pkg, cfg):
    """The north-star bar — MAE <= 1/255, PSNR >= 40 dB — on the low-light scenes needs a reference far beyond 4096 spp
    (two independent 2048-spp oracle halves are only ~28 dB apart).  tests/golden/<cfg>_deep.npz is the oracle's own
    65 536-spp render at quarter resolution (tools/make_goldens.py with GOLDEN_SPP=65536 GOLDEN_DIV=4); the GPU renders
    the same scene instance with 8x as many samples."""
    path = os.path.join(ROOT, "tests", "golden", f"{cfg}_deep.npz")
    if not os.path.exists(path):
        pytest.skip(f"deep golden {cfg} not generated")
    z = np.load(path)
    img, meta = z["img"].astype(np.float64), json.loads(str(z["meta"]))
    spec = pkg.make_scene(meta["scene"], meta["scene_seed"])
    r = pkg.renderer.Renderer(spec, device=0)
    w, h = meta["width"], meta["height"]
    total = np.zeros((h, w, 3), dtype=np.float64)
    spp_total = 0
    for i in range(8):  # 8 launches of golden_spp samples each (distinct seeds = independent sample sets)
        acc, st = r.render(w, h, meta["spp"], meta["depth"], seed=9000 + i, resolve=False)
        total += np.nan_to_num(acc[..., :3].astype(np.float64), nan=0.0, posinf=0.0, neginf=0.0)
        spp_total += meta["spp"]
    gpu = np.sqrt(np.maximum(total / spp_total, 0.0))
    mae = float(np.abs(gpu - img).mean())
    mse = float(((np.clip(gpu, 0, 1) - np.clip(img, 0, 1)) ** 2).mean())
    psnr = float(10 * np.log10(1.0 / max(mse, 1e-12)))
    report = {"config": cfg + "_deep", "scene": meta["scene"], "width": w, "height": h, "golden_spp": meta["spp"],
              "gpu_spp": spp_total, "mae": mae, "psnr_db": psnr, "oracle_half_vs_half_mae": meta["mae_half_vs_half"],
              "oracle_half_vs_half_psnr_db": meta["psnr_half_vs_half"]}
    rp = os.path.join(ROOT, "gpurun_out", "golden_report.json")
    os.makedirs(os.path.dirname(rp), exist_ok=True)
    allr = json.load(open(rp)) if os.path.exists(rp) else {}
    allr[cfg + "_deep"] = report
    json.dump(allr, open(rp, "w"), indent=1)
    print(json.dumps(report))
    assert psnr >= 40.0, report
    assert mae <= 1.0 / 255.0, report


def test_single_process_multi_gpu_render(pkg):
    """hrt_render_multi: one process, N devices, samples sharded, peers' accumulators summed over NVLink peer memory inside
    the resolve kernel.  With one device it must equal hrt_render; with two (when present) it must equal the one-device
    render of the same sample set up to f32 summation order."""
    spec = pkg.make_scene("cornell-smoke", 1)
    r = pkg.renderer.Renderer(spec, device=0)
    w, h, spp = 80, 60, 131
    one, st1 = r.render(w, h, spp, 50, seed=4)
    m1, stm = r.backend.render_multi([0], spec.camera, w, h, spp, 50, spec.background, seed=4)
    assert stm.paths == st1.paths == w * h * spp and stm.rays == st1.rays
    assert np.allclose(np.nan_to_num(m1), np.nan_to_num(one), rtol=2e-4, atol=2e-4)
    n = pkg.native.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs for the peer path")
    devs = list(range(min(n, 8)))
    mN, stN = r.backend.render_multi(devs, spec.camera, w, h, spp, 50, spec.background, seed=4)
    assert stN.paths == st1.paths and stN.rays == st1.rays
    assert np.allclose(np.nan_to_num(mN), np.nan_to_num(one), rtol=3e-4, atol=3e-4)
    acc, _ = r.backend.render_multi(devs, spec.camera, w, h, spp, 50, spec.background, seed=4, resolve=False)
    assert np.all(acc[..., 3] == spp)
