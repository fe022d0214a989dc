"""CPU suite: the oracle (reference sources compiled for the host, oracle/_ref) against the
known-answer vectors the survey pins, and the product's shared host/device headers
(host build, tests/cpu_probe.cpp) against the oracle.  No GPU."""
import ctypes as C

import numpy as np
import pytest

from refhost import RefHost, _p, _f, _up

_fp = C.POINTER(C.c_float)


@pytest.fixture(scope="module")
def ref():
    return RefHost("pt")


def test_rng_known_answer(ref, probe):
    # get_rng(10007,(5,7),(256,256)).state = 0x00511a59, next four states / floats
    probe.probe_rng_seed.restype = C.c_uint32
    want_states = [0x23bdbfe4, 0x64ca89f3, 0x8669c6b6, 0x9cc08e9d]
    want_floats = [0.139614105, 0.393715501, 0.525051534, 0.612313211]
    for seed_fn, draw_fn in ((lambda: ref.lib.ref_rng_seed(10007, 5, 7, 256, 256), ref.lib.ref_rng_draws),
                             (lambda: probe.probe_rng_seed(10007, 5, 7, 256), probe.probe_rng_draws)):
        s = seed_fn()
        assert s == 0x00511a59
        st = np.zeros(4, np.uint32); fl = np.zeros(4, np.float32)
        draw_fn(C.c_uint32(s), 4, _p(st, _up), _p(fl))
        assert list(st) == want_states
        assert np.allclose(fl, want_floats, rtol=0, atol=1e-8)


def test_rng_streams_bit_exact(ref, probe):
    probe.probe_rng_seed.restype = C.c_uint32
    rng = np.random.default_rng(0)
    for _ in range(200):
        f, x, y = int(rng.integers(0, 5000)), int(rng.integers(0, 4096)), int(rng.integers(0, 4096))
        w = int(rng.integers(max(x, 1) + 1, 4097))
        a = ref.lib.ref_rng_seed(f + 10007, x, y, w, 1024)
        b = probe.probe_rng_seed(f + 10007, x, y, w)
        assert a == b
        sa = np.zeros(16, np.uint32); fa = np.zeros(16, np.float32)
        sb = np.zeros(16, np.uint32); fb = np.zeros(16, np.float32)
        ref.lib.ref_rng_draws(C.c_uint32(a), 16, _p(sa, _up), _p(fa))
        probe.probe_rng_draws(C.c_uint32(b), 16, _p(sb, _up), _p(fb))
        assert (sa == sb).all() and (fa.view(np.uint32) == fb.view(np.uint32)).all()


def _unit(rng, n):
    v = rng.standard_normal((n, 3)).astype(np.float32)
    return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


def _hair_inputs(n, seed):
    rng = np.random.default_rng(seed)
    wo, wi, nrm = _unit(rng, n), _unit(rng, n), _unit(rng, n)
    # Y axis of the local frame: unit, h = dot(Y, n) in [-1, 1]
    y = _unit(rng, n)
    h = np.einsum("ij,ij->i", y, nrm).astype(np.float32)
    return wo, wi, nrm, y, h


@pytest.mark.parametrize("beta_m,beta_n,alpha", [(0.3, 0.3, 0.0349065), (0.1, 0.5, 0.0), (0.6, 0.2, 0.05)])
def test_hair_bsdf_eval_matches_reference(ref, probe, beta_m, beta_n, alpha):
    n = 4000
    wo, wi, nrm, y, h = _hair_inputs(n, 1)
    sig = np.array([0.06, 0.1, 0.2], np.float32)
    fa = np.zeros((n, 3), np.float32); pa = np.zeros(n, np.float32)
    fb = np.zeros((n, 3), np.float32); pb = np.zeros(n, np.float32)
    ref.lib.ref_hair_eval(n, _p(wo), _p(wi), _p(nrm), _p(y), _p(sig), C.c_float(beta_m), C.c_float(beta_n), C.c_float(alpha), _p(fa), _p(pa))
    probe.probe_hair_eval(n, _p(wo), _p(wi), _p(h), _p(sig), C.c_float(beta_m), C.c_float(beta_n), C.c_float(alpha), _p(fb), _p(pb))
    # same libm, same operation order: identical up to NaN positions
    assert (np.isnan(fa) == np.isnan(fb)).all()
    ok = ~np.isnan(fa)
    assert np.allclose(fa[ok], fb[ok], rtol=2e-5, atol=1e-7)
    okp = ~np.isnan(pa)
    assert np.allclose(pa[okp], pb[okp], rtol=2e-5, atol=1e-7)
    # tolerance stated by BASELINE.json for fp32 BSDF work: 1e-5 relative on >= 99.9 % of samples
    rel = np.abs(fa[ok] - fb[ok]) / np.maximum(np.abs(fa[ok]), 1e-6)
    assert (rel < 1e-5).mean() > 0.999


def test_hair_bsdf_sample_matches_reference(ref, probe):
    n = 4000
    wo, _, nrm, y, h = _hair_inputs(n, 2)
    u = np.random.default_rng(5).random((n, 4)).astype(np.float32)
    sig = np.array([0.06, 0.1, 0.2], np.float32)
    args = (C.c_float(0.3), C.c_float(0.3), C.c_float(0.0349065))
    wa = np.zeros((n, 3), np.float32); fa = np.zeros((n, 3), np.float32); pa = np.zeros(n, np.float32)
    wb = np.zeros((n, 3), np.float32); fb = np.zeros((n, 3), np.float32); pb = np.zeros(n, np.float32)
    ref.lib.ref_hair_sample(n, _p(wo), _p(nrm), _p(y), _p(u), _p(sig), *args, _p(wa), _p(fa), _p(pa))
    probe.probe_hair_sample(n, _p(wo), _p(h), _p(u), _p(sig), *args, _p(wb), _p(fb), _p(pb))
    ok = ~(np.isnan(wa).any(axis=1) | np.isnan(wb).any(axis=1))
    assert ok.mean() > 0.99
    assert np.allclose(wa[ok], wb[ok], rtol=0, atol=2e-6)
    okf = ok & ~np.isnan(fa).any(axis=1) & ~np.isnan(fb).any(axis=1)
    assert np.allclose(fa[okf], fb[okf], rtol=1e-4, atol=1e-6)
    assert np.allclose(pa[okf], pb[okf], rtol=1e-4, atol=1e-6)


def test_surface_brdf_matches_reference(ref, probe):
    n = 3000
    rng = np.random.default_rng(3)
    wo, wi = _unit(rng, n), _unit(rng, n)
    wo[:, 2] = np.abs(wo[:, 2]); wi[: n // 2, 2] = np.abs(wi[: n // 2, 2])
    kd = np.array([0.3, 0.2, 0.1], np.float32)
    for alpha in (1.0, 0.25):
        fa = np.zeros((n, 3), np.float32); pa = np.zeros(n, np.float32)
        fb = np.zeros((n, 3), np.float32); pb = np.zeros(n, np.float32)
        ref.lib.ref_surf_eval(n, _p(wo), _p(wi), _p(kd), C.c_float(alpha), _p(fa), _p(pa))
        probe.probe_surf_eval(n, _p(wo), _p(wi), _p(kd), C.c_float(alpha), _p(fb), _p(pb))
        assert np.allclose(fa, fb, rtol=1e-6, atol=1e-8, equal_nan=True)
        assert np.allclose(pa, pb, rtol=1e-6, atol=1e-8, equal_nan=True)
        u = rng.random((n, 2)).astype(np.float32)
        wa = np.zeros((n, 3), np.float32); wb = np.zeros((n, 3), np.float32)
        ref.lib.ref_surf_sample(n, _p(wo), _p(u), C.c_float(alpha), _p(wa), _p(pa))
        probe.probe_surf_sample(n, _p(wo), _p(u), C.c_float(alpha), _p(wb), _p(pb))
        assert np.allclose(wa, wb, rtol=0, atol=1e-6, equal_nan=True)
        assert np.allclose(pa, pb, rtol=1e-5, atol=1e-8, equal_nan=True)


def test_curve_hit_geometry_matches_reference(ref, probe):
    rng = np.random.default_rng(4)
    worst = 0.0
    for _ in range(300):
        p0 = rng.uniform(-50, 50, 3)
        step = rng.standard_normal((4, 3)) * 0.5 + rng.standard_normal(3) * 2
        pts = p0 + np.cumsum(step, axis=0)
        cps = np.concatenate([pts, np.full((4, 1), 0.02)], axis=1).astype(np.float32)
        u = np.float32(rng.uniform(0.02, 0.98))
        o = rng.uniform(-200, 200, 3).astype(np.float32)
        # aim at a point near the curve surface
        mid = (cps[1, :3] + cps[2, :3]) / 2
        d = (mid - o); t = np.float32(np.linalg.norm(d) - 0.02); d = (d / np.linalg.norm(d)).astype(np.float32)
        a = np.zeros(13, np.float32); b = np.zeros(13, np.float32)
        ref.lib.ref_curve_geometry(_p(cps), _p(o), _p(d), C.c_float(t), C.c_float(u), _p(a))
        probe.probe_curve_geometry(_p(cps), _p(o), _p(d), C.c_float(t), C.c_float(u), _p(b))
        worst = max(worst, float(np.abs(a - b).max()))
    assert worst < 1e-5


def test_cdf_search_is_lower_bound(probe):
    """The search of the environment CDF rows (hm_light.h: cdf_search) returns the index the reference's
    std::lower_bound order returns (optix_common.cuh:95-149) — incl. flat stretches, the ends and u outside (0,1)."""
    rng = np.random.default_rng(11)
    W, H = 509, 37                                   # w = W + 1 texels per row, like the conditional CDF
    pdf = rng.random((H, W)).astype(np.float32) ** 4
    pdf[:, 100:180] = 0.0                            # flat stretch in every row
    pdf[5] = 0.0; pdf[5, 300] = 1.0                  # a row that is one step
    cdf = np.concatenate([np.zeros((H, 1), np.float32), np.cumsum(pdf, axis=1, dtype=np.float32)], axis=1)
    cdf /= cdf[:, -1:]
    cdf = np.ascontiguousarray(cdf, np.float32)
    n = 20000
    u = rng.random(n).astype(np.float32)
    u[:50] = 0.0; u[50:100] = 1.0; u[100:150] = -0.5; u[150:200] = 1.5
    u[200:400] = cdf[rng.integers(0, H, 200), rng.integers(0, W + 1, 200)]     # exactly on table values
    yn = ((rng.integers(0, H, n) + 0.5) / H).astype(np.float32)
    out = np.zeros((n, 2), np.int32)
    probe.probe_cdf_search(_p(cdf), W + 1, H, n, _p(u), _p(yn), C.c_float(W), out.ctypes.data_as(C.POINTER(C.c_int)))
    assert np.array_equal(out[:, 0], out[:, 1])
    rows = np.floor(yn * H).astype(int)
    want = np.array([max(int(np.searchsorted(cdf[r, :W], uu, side="left")) - 1, 0) for r, uu in zip(rows, u)])
    assert np.array_equal(out[:, 0], want)
    assert out.min() >= 0 and out.max() <= W - 1 and len(np.unique(out[:, 0])) > 200
    # marginal-CDF shape: one row
    m = np.ascontiguousarray(cdf[7])
    out1 = np.zeros((n, 2), np.int32)
    probe.probe_cdf_search(_p(m), W + 1, 1, n, _p(u), _p(np.zeros(n, np.float32)), C.c_float(W), out1.ctypes.data_as(C.POINTER(C.c_int)))
    assert np.array_equal(out1[:, 0], out1[:, 1])


@pytest.mark.parametrize("W", [64, 256, 4096])
def test_two_level_cdf_search_equals_reference_order_search(probe, W):
    """hm_light.h: cdf_lower_bound_two_level (what the device uses on power-of-two maps) returns the index of the
    reference-order search (normalised-coordinate table fetches, std::lower_bound probe order) for every u — flat
    stretches, single-step rows, values exactly on table entries, u outside (0, 1)."""
    rng = np.random.default_rng(W)
    H = 9
    pdf = rng.random((H, W)).astype(np.float32) ** 4
    pdf[:, W // 4:W // 2] = 0.0
    pdf[3] = 0.0; pdf[3, W - 3] = 1.0
    pdf[4] = 0.0; pdf[4, 0] = 1.0
    cdf = np.concatenate([np.zeros((H, 1), np.float32), np.cumsum(pdf, axis=1, dtype=np.float32)], axis=1)
    cdf = np.ascontiguousarray(cdf / cdf[:, -1:], np.float32)
    n = 30000
    u = rng.random(n).astype(np.float32)
    u[:50] = 0.0; u[50:100] = 1.0; u[100:150] = -0.5; u[150:200] = 1.5
    rows = rng.integers(0, H, n).astype(np.int32)
    u[200:2200] = cdf[rows[200:2200], rng.integers(0, W + 1, 2000)]
    out = np.zeros((n, 2), np.int32)
    probe.probe_cdf_two_level(_p(cdf), W, H, n, _p(u), rows.ctypes.data_as(C.POINTER(C.c_int)), out.ctypes.data_as(C.POINTER(C.c_int)))
    assert np.array_equal(out[:, 0], out[:, 1])
    assert len(np.unique(out[:, 0])) > min(W // 2, 200)
