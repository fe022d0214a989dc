"""Multi-GPU product path: NCCL communicator inside libhairmsnn.so — gradient all-reduce in the
frame loop, framebuffer reduce — with one process per GPU.  Needs >= 2 devices (gpurun --gpus 2); the single-GPU
reference renders happen in this process on device 0."""
import os
import subprocess
import sys

import numpy as np
import pytest

from common import small_scene_kwargs

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
WORLD = 2
FRAMES = 4


def _need_two_gpus():
    from hairmsnn_b200 import api
    n = api.device_count()
    assert n > 0, "no CUDA device"
    if n < WORLD:
        pytest.skip(f"needs {WORLD} GPUs, this box has {n} (run with gpurun --gpus 2)")


def run_ranks(tmp_path, mode, frames=FRAMES):
    id_file = str(tmp_path / f"{mode}.id")
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "multirank_worker.py"), str(r), str(WORLD), id_file, str(tmp_path), mode, str(frames)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(WORLD)]
    outs = []
    for p in procs:
        o, _ = p.communicate(timeout=600)
        outs.append(o)
    for r, p in enumerate(procs):
        assert p.returncode == 0, f"rank {r} failed:\n{outs[r][-3000:]}"
    return [np.load(tmp_path / f"{mode}_rank{r}.npz") for r in range(WORLD)]


def single_gpu(kind, frames, pretrain=0):
    from hairmsnn_b200 import api
    kw = small_scene_kwargs(width=128, height=128, strands=800, segs=12, path_v2=8)
    sc = api.Scene.from_arrays(**kw)
    r = api.Renderer(sc, kind, beta_cli=1, device=0)
    if pretrain:
        r.msnn_pretrain(pretrain)
    r.render_frames(frames)
    out = {"final_avg": r.buffer(api.BUF_FINAL_AVG), "fb8": r.buffer(api.BUF_FB8), "final_accum": r.buffer(api.BUF_FINAL_ACCUM)}
    if kind == api.HAIR_MSNN:
        out["pt_avg"] = r.buffer(api.BUF_PT_AVG)
    r.close()
    return out


def test_pt_sample_groups_sum_to_single_gpu(tmp_path):
    """2 ranks x FRAMES samples (ids r, r+2, ...) reduced == one GPU rendering ids 0..2*FRAMES-1 (float sum order differs)."""
    _need_two_gpus()
    from hairmsnn_b200 import api
    res = run_ranks(tmp_path, "pt_spp")
    ref = single_gpu(api.PATH_TRACING, WORLD * FRAMES)
    a, b = res[0]["final_avg"], res[1]["final_avg"]
    assert np.array_equal(a, b), "ranks hold different reduced images"
    # the sum of the ranks' local accumulation buffers is the single-GPU accumulation up to fp32 summation order
    local_sum = res[0]["local_final_accum"][..., :3] + res[1]["local_final_accum"][..., :3]
    np.testing.assert_allclose(local_sum, ref["final_accum"][..., :3], rtol=2e-5, atol=1e-5)
    np.testing.assert_allclose(a[..., :3], ref["final_avg"][..., :3], rtol=2e-5, atol=1e-5)
    assert (np.abs(a[..., :3]).sum() > 0)
    # 8-bit frames agree except where the float difference crosses a quantisation step
    assert (res[0]["fb8"] != ref["fb8"]).mean() < 1e-3


def test_pt_row_bands_bit_exact(tmp_path):
    """Row bands reproduce the single-GPU frame bit-for-bit (RNG keyed by full-frame pixel index)."""
    _need_two_gpus()
    from hairmsnn_b200 import api
    res = run_ranks(tmp_path, "pt_bands")
    ref = single_gpu(api.PATH_TRACING, FRAMES)
    for r in range(WORLD):
        assert np.array_equal(res[r]["final_avg"], ref["final_avg"])
        assert np.array_equal(res[r]["fb8"], ref["fb8"])
    # a rank's own accumulation buffer is zero outside its band
    H = ref["final_avg"].shape[0]
    assert not res[0]["local_final_accum"][H // 2:].any() and not res[1]["local_final_accum"][:H // 2].any()


@pytest.mark.parametrize("mode", ["msnn_spp", "msnn_bands", "nrc_spp"])
def test_training_replicas_stay_identical(tmp_path, mode):
    """Gradient all-reduce inside the frame loop: after pre-training + FRAMES steps every rank holds bit-identical
    weights that moved away from the initial ones; the reduced images are identical on all ranks and finite."""
    _need_two_gpus()
    from hairmsnn_b200 import api
    res = run_ranks(tmp_path, mode)
    p0, p1 = res[0]["params"], res[1]["params"]
    assert np.array_equal(p0, p1), f"replicas diverged: {np.abs(p0 - p1).max()}"
    m = api.Mlp.create(None, 9 if mode.startswith("nrc") else 12, 3, 0)
    init = m.get_params()
    m.close()
    assert np.abs(p0 - init).max() > 1e-4, "weights did not train"
    assert np.array_equal(res[0]["final_avg"], res[1]["final_avg"])
    assert np.isfinite(res[0]["final_avg"]).all()
    if mode == "msnn_bands":
        # the path-traced component does not depend on the network: bit-exact against one GPU
        ref = single_gpu(api.HAIR_MSNN, FRAMES, pretrain=3)
        assert np.array_equal(res[0]["pt_avg"], ref["pt_avg"])
        # the composite differs only through the network (training order of atomics, batch split)
        d = np.abs(res[0]["final_avg"][..., :3] - ref["final_avg"][..., :3])
        assert d.mean() < 0.05 * max(ref["final_avg"][..., :3].mean(), 1e-3)


def test_executables_with_gpus_flag(tmp_path):
    """`render_path_tracing <config> --gpus 2 --shard bands` writes the image a single GPU writes, bit for bit; `render_hair_msnn
    <config> 1 --gpus 2` (sample groups, gradients all-reduced inside the loop) writes a finite image of the same scene."""
    _need_two_gpus()
    import json
    from hairmsnn_b200 import api
    root = os.path.dirname(HERE)
    curly = os.path.join(root, "assets", "scenes", "curly", "config.json")
    if not os.path.exists(curly):
        pytest.skip("assets/scenes is not staged")
    cfg = json.load(open(curly))
    cfg["integrator"]["width"] = cfg["integrator"]["height"] = 256
    path = os.path.join(os.path.dirname(curly), "config_test_256.json")
    json.dump(cfg, open(path, "w"))
    env = dict(os.environ, HM_BVH_CACHE=str(tmp_path))
    exe = os.path.join(root, "hairmsnn_b200", "bin")

    def run(name, *args):
        out = str(tmp_path / (name + ".png"))
        p = subprocess.run(list(args) + ["--out", out], capture_output=True, text=True, timeout=900, env=env)
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
        return api.load_exr(out.replace(".png", ".exr")), p.stdout
    one, _ = run("pt1", os.path.join(exe, "render_path_tracing"), path, "--spp", "4")
    two, log = run("pt2", os.path.join(exe, "render_path_tracing"), path, "--spp", "4", "--gpus", "2", "--shard", "bands")
    assert "on 2 GPU(s) (row bands)" in log
    assert np.array_equal(one, two)
    img, log = run("msnn2", os.path.join(exe, "render_hair_msnn"), path, "1", "--spp", "8", "--gpus", "2", "--pretrain-steps", "20")
    assert "on 2 GPU(s) (sample groups)" in log
    assert img.shape == (256, 256, 4) and np.isfinite(img).all() and img[..., :3].mean() > 0.01
    assert np.abs(img[..., :3].mean() - one[..., :3].mean()) < 0.2 * one[..., :3].mean()
