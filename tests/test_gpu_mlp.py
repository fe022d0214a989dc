"""GPU parity: the fused encode+MLP kernels (through the C ABI) against the numpy restatement
of the tiny-cuda-nn configuration (oracle/mlp_oracle.py).

Tolerances (BASELINE.json north_star: "MLP outputs within a stated fp32/fp16 tolerance"):
  * initial parameters: bit-exact (same PRNG streams)
  * inference vs the fp16-storage oracle : |dy| <= 2e-3 * max(1, |y|)   (we accumulate in fp32; the
    reference accumulates in fp16 inside wmma, so its own error vs this oracle is larger)
  * inference vs the fp32 oracle         : |dy| <= 2e-2 * max(1, |y|)
  * parameter gradients vs the fp16-storage oracle: relative L2 error <= 2e-2 per matrix / grid
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from hairmsnn_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mlp_oracle as mo  # noqa: E402

pytestmark = pytest.mark.gpu


def _cudart():
    for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
        try:
            return C.CDLL(name)
        except OSError:
            continue
    raise RuntimeError("libcudart not found")


def _d2h(ptr, nbytes):
    out = np.empty(nbytes, np.uint8)
    rt = _cudart()
    rt.cudaDeviceSynchronize()
    assert rt.cudaMemcpy(out.ctypes.data_as(C.c_void_p), C.c_void_p(ptr), C.c_size_t(nbytes), 2) == 0
    return out


def _inputs(n, seed, in_ch=12):
    rng = np.random.default_rng(seed)
    d = rng.standard_normal((n, 6)); d[:, :3] /= np.linalg.norm(d[:, :3], axis=1, keepdims=True); d[:, 3:] /= np.linalg.norm(d[:, 3:], axis=1, keepdims=True)
    x = np.concatenate([rng.uniform(-0.5, 0.5, (n, 3)), d, np.zeros((n, in_ch - 9))], axis=1).astype(np.float32)
    return x


def _trained_like_params(cfg, seed=3):
    """Initial params have a ~1e-4 grid (everything rounds the same); make the test bite."""
    rng = np.random.default_rng(seed)
    p = mo.initial_params(cfg)
    p[9216:] = rng.uniform(-0.3, 0.3, p.size - 9216).astype(np.float32)
    return p


def test_initial_parameters_bit_exact():
    m = api.Mlp.create()
    cfg = mo.Config(12)
    assert m.n_params == mo.n_params(cfg)[0] == 1000448
    got = m.get_params()
    want = mo.initial_params(cfg)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("in_ch", [12, 9])
def test_inference_matches_oracle(in_ch):
    cfg = mo.Config(in_ch)
    m = api.Mlp.create(in_ch=in_ch)
    p = _trained_like_params(cfg)
    m.set_params(p)
    x = _inputs(1024, 0, in_ch)
    y = m.inference(x)
    y16 = mo.forward(cfg, p, x, half=True)
    y32 = mo.forward(cfg, p, x, half=False)
    assert np.abs(y32).max() > 0.05
    assert (np.abs(y - y16) <= 2e-3 * np.maximum(1, np.abs(y16))).all(), np.abs(y - y16).max()
    assert (np.abs(y - y32) <= 2e-2 * np.maximum(1, np.abs(y32))).all()


def test_inference_edge_cases():
    m = api.Mlp.create()
    cfg = mo.Config(12)
    p = _trained_like_params(cfg)
    m.set_params(p)
    # negative coordinates wrap in the dense levels, directions outside [0,1], exact cell corners
    x = _inputs(128, 1)
    x[:16, :3] = -0.49; x[16:32, :3] = 0.0; x[32:48, :3] = 1.0 / 15.0; x[48:64, 3:9] = -1.0; x[64:80, 3:9] = 1.0
    y = m.inference(x)
    y16 = mo.forward(cfg, p, x, half=True)
    assert (np.abs(y - y16) <= 4e-3 * np.maximum(1, np.abs(y16))).all()
    with pytest.raises(api.HairMSNNError):
        m.inference(x[:100])           # not a multiple of 128 (tcnn batch granularity)
    m.reset()                          # TINY_MLP::reset -> all-zero weights -> zero output
    assert np.all(m.inference(x) == 0)


def test_gradients_match_oracle():
    cfg = mo.Config(12)
    m = api.Mlp.create()
    p = _trained_like_params(cfg)
    m.set_params(p)
    n = 2048
    x = _inputs(n, 2)
    y = np.random.default_rng(5).uniform(0, 1, (n, 3)).astype(np.float32)
    rt = _cudart()
    d_x, d_y = C.c_void_p(), C.c_void_p()
    rt.cudaMalloc(C.byref(d_x), x.nbytes); rt.cudaMalloc(C.byref(d_y), y.nbytes)
    rt.cudaMemcpy(d_x, x.ctypes.data_as(C.c_void_p), C.c_size_t(x.nbytes), 1)
    rt.cudaMemcpy(d_y, y.ctypes.data_as(C.c_void_p), C.c_size_t(y.nbytes), 1)
    m.forward_backward_device(d_x.value, d_y.value, n)
    gptr, gcount = m.gradients_device()
    g = _d2h(gptr, gcount * 4).view(np.float32)
    loss = m.loss()
    want_loss, want = mo.backward(cfg, p, x, y, half=True)
    assert abs(loss - want_loss) <= 2e-3 * abs(want_loss)
    for name, a, b in (("W0", 0, 4096), ("W1", 4096, 8192), ("Wout", 8192, 9216), ("grid", 9216, g.size)):
        num = np.linalg.norm(g[a:b] - want[a:b]); den = np.linalg.norm(want[a:b])
        assert den > 0 and num / den <= 2e-2, (name, num / den)
    # rows 3..15 of the padded output matrix receive no gradient
    assert np.all(g[8192 + 3 * 64:9216] == 0)
    rt.cudaFree(d_x); rt.cudaFree(d_y)


def test_training_step_follows_adam_oracle_and_learns():
    cfg = mo.Config(12)
    m = api.Mlp.create()
    p0 = m.get_params()
    n = 1024
    x = _inputs(n, 7)
    tgt = (0.5 + 0.5 * np.sin(4 * x[:, :3])).astype(np.float32)
    loss0 = m.train_step(x, tgt)
    p1 = m.get_params()
    # oracle: same step from the same start
    want_loss, g = mo.backward(cfg, p0, x, tgt, half=True)
    opt = mo.Adam(cfg, p0)
    want_p1 = opt.step(g, half=True)
    assert abs(loss0 - want_loss) <= 5e-3 * abs(want_loss)
    moved = np.abs(p1 - p0) > 0
    want_moved = np.abs(want_p1 - p0) > 0
    # first Adam step is +-lr for every parameter with a non-zero (fp16) gradient
    agree = (moved == want_moved).mean()
    assert agree > 0.995, agree
    both = moved & want_moved
    assert np.allclose(p1[both], want_p1[both], atol=2e-3)
    assert (np.sign(p1[both] - p0[both]) == np.sign(want_p1[both] - p0[both])).mean() > 0.995
    losses = [loss0] + [m.train_step(x, tgt) for _ in range(60)]
    assert losses[-1] < 0.5 * losses[0], losses[::10]
    m.reinitialize()
    assert np.array_equal(m.get_params(), p0)


def test_save_and_load_roundtrip(tmp_path):
    m = api.Mlp.create()
    x = _inputs(128, 9)
    m.train_step(x, np.full((128, 3), 0.3, np.float32))
    p = m.get_params()
    path = str(tmp_path / "w.bin")
    m.save(path)
    m2 = api.Mlp.create()
    m2.load(path)
    assert np.array_equal(m2.get_params(), p)
    assert np.array_equal(m2.inference(x), m.inference(x))
    with pytest.raises(api.HairMSNNError):
        m2.load(str(tmp_path / "missing.bin"))


def test_tcnn_snapshot_roundtrip_and_half_snapshots(tmp_path):
    """Trainer::serialize / deserialize (trainer.h:270-310) in the text-JSON form TINY_MLP::loadWeights reads."""
    import json
    a = api.Mlp.create()
    x = np.random.default_rng(3).uniform(-0.5, 0.5, (256, 12)).astype(np.float32)
    y = np.random.default_rng(4).uniform(0, 1, (256, 3)).astype(np.float32)
    for _ in range(3):
        a.train_step(x, y)
    pa = a.get_params()
    p = str(tmp_path / "snap.json")
    a.save_snapshot(p)
    j = json.load(open(p))
    assert j["n_params"] == 1000448 and j["params_type"] == "float" and len(j["params_binary"]["bytes"]) == 4 * 1000448
    b = api.Mlp.create()
    assert not np.array_equal(b.get_params(), pa)
    b.load(p)
    assert np.array_equal(b.get_params(), pa)
    assert np.array_equal(b.inference(x), a.inference(x))
    # a snapshot as the reference build writes it: params_type "__half"
    half = pa.astype(np.float16)
    j2 = {"n_params": int(pa.size), "params_type": "__half", "params_binary": {"bytes": half.view(np.uint8).tolist(), "subtype": None}}
    p2 = str(tmp_path / "snap_half.json")
    json.dump(j2, open(p2, "w"))
    c = api.Mlp.create()
    c.load(p2)
    assert np.array_equal(c.get_params(), half.astype(np.float32))
    # wrong size / wrong type are errors
    j2["params_binary"]["bytes"] = j2["params_binary"]["bytes"][:-2]
    json.dump(j2, open(p2, "w"))
    with pytest.raises(api.HairMSNNError):
        c.load(p2)
    j2["params_type"] = "double"
    json.dump(j2, open(p2, "w"))
    with pytest.raises(api.HairMSNNError):
        c.load(p2)


# ---- pinned against the reference's own tiny-cuda-nn (tests/golden/tcnn_{12,9}.npz) -----------------------------
# The vectors are outputs of /root/reference/extern/tiny-cuda-nn built for sm_100 and driven as TINY_MLP does
# (oracle/tcnn_golden.cu), one run on a B200.  Tolerances are the fp16-accumulation ones SURVEY §7 suggests:
# tcnn accumulates in fp16 inside wmma and scatters grid gradients with half2 atomics; these kernels accumulate
# in fp32 and round where tcnn stores fp16.
import tcnn_inputs  # noqa: E402


def _rl2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


@pytest.mark.parametrize("in_ch", [12, 9])
def test_against_tiny_cuda_nn_golden_vectors(in_ch):
    g = np.load(os.path.join(ROOT, "tests", "golden", f"tcnn_{in_ch}.npz"))
    n, rows, stride = int(g["n_rows"]), int(g["rows_kept"]), int(g["grid_stride"])
    x = tcnn_inputs.make_inputs(n, in_ch, int(g["input_seed"]))
    y = tcnn_inputs.make_targets(n, int(g["target_seed"]))
    m = api.Mlp.create(in_ch=in_ch)
    # initial parameters: bit-exact (Trainer::initialize_params incl. the FMA of the grid initialisation)
    p0 = m.get_params()
    assert p0.size == int(g["n_params"])
    assert np.array_equal(p0[:9216].view(np.uint32), g["params0_f32.mlp"].view(np.uint32))
    assert np.array_equal(p0[9216::stride].view(np.uint32), g["params0_f32.grid_sample"].view(np.uint32))
    assert abs(p0.astype(np.float64).sum() - float(g["params0_f32.sum"])) < 1e-12
    # TINY_MLP::inference
    out = m.inference(x)
    assert np.abs(out[:rows] - g["infer0"]).max() <= 1e-3, np.abs(out[:rows] - g["infer0"]).max()
    assert np.allclose(out.astype(np.float64).sum(axis=0), g["infer0.sum"], rtol=2e-3, atol=0.5)
    # Trainer::training_step x 4 on one batch
    rt = _cudart()
    d_x, d_y = C.c_void_p(), C.c_void_p()
    rt.cudaMalloc(C.byref(d_x), x.nbytes); rt.cudaMalloc(C.byref(d_y), y.nbytes)
    rt.cudaMemcpy(d_x, x.ctypes.data_as(C.c_void_p), C.c_size_t(x.nbytes), 1)
    rt.cudaMemcpy(d_y, y.ctypes.data_as(C.c_void_p), C.c_size_t(y.nbytes), 1)
    for s in range(4):
        m.forward_backward_device(d_x.value, d_y.value, n)
        loss = m.loss()
        assert abs(loss - float(g["losses"][s])) <= 1e-3 * float(g["losses"][s]), (s, loss, float(g["losses"][s]))
        if s == 0:
            gptr, gcount = m.gradients_device()
            gr = _d2h(gptr, gcount * 4).view(np.float32)
            assert _rl2(gr[:9216], g["grad1_f16.mlp"].astype(np.float32)) < 5e-3
            assert _rl2(gr[9216::stride], g["grad1_f16.grid_sample"].astype(np.float32)) < 3e-2
        m.optimizer_step()
        if s == 0:
            p1 = m.get_params()
            assert _rl2(p1[:9216], g["params1_f32.mlp"]) < 2e-2
            moved_ref = g["params1_f32.grid_sample"] != g["params0_f32.grid_sample"]
            assert ((p1[9216::stride] != p0[9216::stride]) == moved_ref).mean() > 0.999
    p4 = m.get_params()
    assert _rl2(p4[:9216], g["params4_f32.mlp"]) < 2e-2
    assert _rl2(p4[9216::stride], g["params4_f32.grid_sample"]) < 0.1
    out4 = m.inference(x)
    assert _rl2(out4[:rows], g["infer4"]) < 5e-3
    # TINY_MLP::reset(): weights zeroed, optimiser state kept; one more step from there
    m.reset()
    m.forward_backward_device(d_x.value, d_y.value, n)
    assert abs(m.loss() - float(g["loss_reset"][0])) <= 1e-3 * float(g["loss_reset"][0])
    rt.cudaFree(d_x); rt.cudaFree(d_y)
