"""GPU tests on the reference's shipped scene (scenes/curly: real .hair geometry, head mesh, PIZ environment map),
loaded through hm_scene_load from the staged copy under assets/scenes (scripts/stage_assets.py), at a reduced frame
size and sample count — and of the three headless executables on it."""
import json
import os
import struct
import subprocess

import numpy as np
import pytest

from hairmsnn_b200 import api
from common import camera_rays
from refhost import RefHost

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CURLY = os.path.join(ROOT, "assets", "scenes", "curly", "config.json")
BIN = os.path.join(ROOT, "hairmsnn_b200", "bin")


def _need_assets():
    if not os.path.exists(CURLY):
        pytest.skip("assets/scenes is not staged (python scripts/stage_assets.py where /root/reference exists)")


def _small_config(size=256):
    """scenes/curly/config.json with a smaller frame, written next to it (paths inside resolve relative to its directory)."""
    cfg = json.load(open(CURLY))
    cfg["integrator"]["width"] = cfg["integrator"]["height"] = size
    out = os.path.join(os.path.dirname(CURLY), f"config_test_{size}.json")
    json.dump(cfg, open(out, "w"))
    return out, cfg


@pytest.fixture(scope="module")
def curly(tmp_path_factory):
    _need_assets()
    cache = tmp_path_factory.mktemp("bvh_cache")
    path, cfg = _small_config(256)
    os.environ["HM_BVH_CACHE"] = str(cache)      # the executables started below restore the tree instead of rebuilding it
    sc = api.Scene.load(path)
    yield sc, path, cfg
    os.environ.pop("HM_BVH_CACHE", None)


def test_shipped_curly_scene_loads(curly):
    sc, _, _ = curly
    i = sc.info()
    assert (i.num_segments, i.num_strands, i.num_triangles) == (3391580, 50000, 78520)
    assert (i.env_w, i.env_h, i.num_dlights, i.spp, i.path_v2) == (4096, 2048, 1, 500, 40)
    assert abs(i.scene_scale - 187.85) < 0.5


def test_env_tables_of_the_shipped_map_built_on_device(curly):
    """The 4096 x 2048 studio map: the device-built importance tables equal the host recipe's, entry for entry."""
    sc, _, _ = curly
    r = api.Renderer(sc, api.PATH_TRACING)
    t = sc.env_tables()
    for which, name in ((api.BUF_ENV_CPDF, "cpdf"), (api.BUF_ENV_CCDF, "ccdf"), (api.BUF_ENV_MPDF, "mpdf"), (api.BUF_ENV_MCDF, "mcdf")):
        got = r.buffer(which)
        assert np.array_equal(got.view(np.uint32), np.ascontiguousarray(t[name]).reshape(-1).view(np.uint32)), name


def test_hit_ids_bit_exact_on_real_hair(curly):
    """Primary-hit curve / segment ids, t and u on the real geometry: the sm_100a traversal of the 8-wide tree == the host
    traversal of the binary tree it is derived from (same intersector source), for camera rays and for incoherent rays."""
    sc, _, _ = curly
    r = api.Renderer(sc, api.PATH_TRACING)
    ref = RefHost("pt")
    ref.bind_scene(sc)
    o, d = camera_rays(sc.info(), 20000, seed=3)
    rng = np.random.default_rng(4)
    o2 = rng.uniform(-60, 60, (20000, 3)).astype(np.float32)
    d2 = rng.standard_normal((20000, 3)).astype(np.float32); d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
    for org, dirs in ((o, d), (o2, d2)):
        g = r.trace_rays(org, dirs)
        t, p, u = ref.trace_ids(org, dirs)
        assert (p >= 0).mean() > 0.1
        assert np.array_equal(g["prim"], p)
        hit = p >= 0
        assert np.array_equal(g["t"][hit].view(np.uint32), t[hit].view(np.uint32))
        assert np.array_equal(g["u"][hit].view(np.uint32), u[hit].view(np.uint32))
    ga = r.trace_rays(o2, d2, any_hit=True)
    _, pa, _ = ref.trace_ids(o2, d2, any_hit=True)
    assert np.array_equal(ga["prim"] >= 0, pa >= 0)


def test_reduced_render_matches_reference_and_cache_tracks_path_tracer(curly):
    sc, path, cfg = curly
    W = H = 256
    pt = api.Renderer(sc, api.PATH_TRACING)
    pt.render_frames(1)
    # frame 0 against the reference's rayGenCam on the host (rows through the hair volume)
    from bench import kw_from_config
    ref = RefHost("pt")
    ref.bind_all(sc, kw_from_config(path))
    y0, y1 = 120, 136
    accum, _, _ = ref.render_pt(0, W, H, y0=y0, y1=y1)
    g = pt.buffer(api.BUF_FINAL_ACCUM)[y0:y1, :, :3]; w = accum[y0:y1, :, :3]
    err = np.abs(g - w).max(axis=2) / np.maximum(np.abs(w).max(axis=2), 1e-2)
    # 40-vertex paths through real hair: every vertex is a chance for a libdevice/glibc ulp to flip a discrete choice
    # (see test_gpu_pt.py on the tolerance); the synthetic 128^2 scenes reach 96 %, this one ~91 %
    assert (err < 2e-3).mean() > 0.88, (err < 2e-3).mean()
    assert abs(g.mean() - w.mean()) < 0.05 * w.mean()
    pt.render_frames(63)
    truth = pt.buffer(api.BUF_FINAL_AVG)[..., :3]
    m = api.Renderer(sc, api.HAIR_MSNN, beta_cli=1)
    m.msnn_pretrain(100)
    m.render_frames(64)
    final = m.buffer(api.BUF_FINAL_AVG)[..., :3]
    assert np.isfinite(final).all()
    flags = m.buffer(api.BUF_GBUFFER).reshape(H, W, 4)[..., 3].copy().view(np.int32)
    hair = ((flags & 1) != 0) & ((flags & 2) == 0)
    assert 0.3 < hair.mean() < 0.8
    # relMSE as bench.py's image gate defines it, here at 64 spp / 256^2 (noisy truth): loose bound
    rel = float(np.mean((final - truth) ** 2 / (truth ** 2 + 1e-2)))
    assert rel < 0.2, rel
    # after 100 + 64 training steps the cache has recovered a good part of the energy the truncated paths lose
    short = m.buffer(api.BUF_PT_AVG)[..., :3]
    e_short = abs(float(short[hair].mean()) - float(truth[hair].mean()))
    e_final = abs(float(final[hair].mean()) - float(truth[hair].mean()))
    assert short[hair].mean() < truth[hair].mean() and e_final < 0.8 * e_short, (e_short, e_final)


def _png_size(path):
    b = open(path, "rb").read(32)
    assert b[:8] == b"\x89PNG\r\n\x1a\n"
    return struct.unpack(">II", b[16:24])


@pytest.mark.parametrize("exe,args", [("render_hair_msnn", ["1", "--pretrain-steps", "20"]), ("render_path_tracing", []), ("render_nrc", [])])
def test_executables_write_png_exr_stats(curly, tmp_path, exe, args):
    """`exe <config.json> [BETA] --spp 4` (render_hair_msnn.cu:1146-1175 main): PNG + EXR (+ _pt / _nn components) + stats."""
    _, path, _ = curly
    out = str(tmp_path / "render.png")
    p = subprocess.run([os.path.join(BIN, exe), path] + args + ["--spp", "4", "--out", out], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "Mpaths/s" in p.stdout
    assert _png_size(out) == (256, 256)
    img = api.load_exr(out.replace(".png", ".exr"))
    assert img.shape == (256, 256, 4) and np.isfinite(img).all() and img[..., :3].mean() > 0.01
    if exe == "render_hair_msnn":
        a, b = api.load_exr(out.replace(".png", "_pt.exr")), api.load_exr(out.replace(".png", "_nn.exr"))
        assert np.isfinite(a).all() and np.isfinite(b).all()
    st = json.load(open(out.replace(".png", "_stats.json")))
    assert st and isinstance(st, dict)
    # the frame spans cover the job: 4 frames of 256 x 256 paths, at a rate a B200 can have
    assert st["spp"] == 4 and st["paths"] == 4 * 256 * 256
    assert st["seconds"] > 0 and 0.1 < st["mpaths_per_s"] < 2000, (st["seconds"], st["mpaths_per_s"])


def test_executable_reports_scene_errors(tmp_path):
    p = subprocess.run([os.path.join(BIN, "render_path_tracing"), str(tmp_path / "missing.json")], capture_output=True, text=True, timeout=60)
    assert p.returncode != 0 and "Error loading scene" in p.stderr
