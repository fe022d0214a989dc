"""GPU parity of the fibre scattering model: the sm_100a build of hm_bsdf.h through the C ABI
(hm_bsdf_eval / hm_bsdf_sample) against the REFERENCE's disney_hair.cuh compiled for the host (oracle/_ref).

Tolerance (BASELINE.json north_star: "BSDF eval/sample within a stated fp32 tolerance"): libdevice's expf / logf / sinhf /
atan2f / asinf differ from glibc's by a few ulp, and the model chains several of them (exp of a log-Bessel, sinh, logistic):
|df| <= 2e-4 * max(|f|, 1e-3) on >= 99.5 % of samples and <= 5e-3 relative on all finite ones; NaNs in the same places."""
import ctypes as C

import numpy as np
import pytest

from hairmsnn_b200 import api
from refhost import RefHost, _p

pytestmark = pytest.mark.gpu


def _unit(rng, n):
    v = rng.standard_normal((n, 3)).astype(np.float32)
    return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


def _inputs(n, seed):
    rng = np.random.default_rng(seed)
    wo, wi, nrm, y = _unit(rng, n), _unit(rng, n), _unit(rng, n), _unit(rng, n)
    # edge cases the reference's tests of this model would want: h exactly +-1 (gamma_o = +-pi/2), h = 0,
    # grazing and axial outgoing directions
    nrm[:64] = y[:64]; nrm[64:128] = -y[64:128]
    nrm[128:192] = np.cross(y[128:192], _unit(rng, 64)); nrm[128:192] /= np.linalg.norm(nrm[128:192], axis=1, keepdims=True)
    wo[192:224] = np.array([0.9999, 0.01, 0.0], np.float32); wo[224:256] = np.array([0.0, 0.6, 0.8], np.float32)
    wo[192:256] /= np.linalg.norm(wo[192:256], axis=1, keepdims=True)
    h = np.einsum("ij,ij->i", y, nrm).astype(np.float32)
    return wo, wi, nrm.astype(np.float32), y, h


def _close(got, want, name):
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    assert (nan_g == nan_w).mean() > 0.999, name
    ok = ~(nan_g | nan_w) & np.isfinite(want)
    rel = np.abs(got[ok] - want[ok]) / np.maximum(np.abs(want[ok]), 1e-3)
    assert (rel <= 2e-4).mean() >= 0.995, (name, (rel <= 2e-4).mean(), rel.max())
    assert rel.max() <= 5e-3, (name, rel.max())


# beta_m = 0.3 -> v = (0.0846, 0.0212, 0.339): both branches of Mp (v <= 0.1 and v > 0.1); 0.1 -> all three below; 0.6 -> all above
@pytest.mark.parametrize("beta_m,beta_n,alpha", [(0.3, 0.3, 0.0349065), (0.1, 0.5, 0.0), (0.6, 0.2, 0.05)])
def test_hair_eval_on_device_matches_reference_header(beta_m, beta_n, alpha):
    n = 8192
    wo, wi, nrm, y, h = _inputs(n, 1)
    sig = np.array([0.06, 0.1, 0.2], np.float32)
    ref = RefHost("pt")
    fa = np.zeros((n, 3), np.float32); pa = np.zeros(n, np.float32)
    ref.lib.ref_hair_eval(n, _p(wo), _p(wi), _p(nrm), _p(y), _p(sig), C.c_float(beta_m), C.c_float(beta_n), C.c_float(alpha), _p(fa), _p(pa))
    f, pdf = api.bsdf_eval(wo, wi, h, sig, beta_m, beta_n, alpha)
    assert np.isfinite(fa).mean() > 0.9 and np.abs(fa[np.isfinite(fa)]).max() > 0.1
    _close(f, fa, "f")
    _close(pdf, pa, "pdf")


def test_hair_sample_on_device_matches_reference_header():
    n = 8192
    wo, _, nrm, y, h = _inputs(n, 2)
    rng = np.random.default_rng(5)
    u = rng.random((n, 4)).astype(np.float32)
    u[:512, 0] = 0.999 + 0.001 * rng.random(512).astype(np.float32)     # the residual lobe (p = 3: sampled with v[3] = s)
    u[512:640, 1] = 0.0                                                  # eps2 clamp at 1e-5
    u[640:700, 3] = 0.0; u[700:760, 3] = np.float32(0.99999994)
    sig = np.array([0.06, 0.1, 0.2], np.float32)
    ref = RefHost("pt")
    wa = np.zeros((n, 3), np.float32); fa = np.zeros((n, 3), np.float32); pa = np.zeros(n, np.float32)
    ref.lib.ref_hair_sample(n, _p(wo), _p(nrm), _p(y), _p(u), _p(sig), C.c_float(0.3), C.c_float(0.3), C.c_float(0.0349065), _p(wa), _p(fa), _p(pa))
    wi, f, pdf = api.bsdf_sample(wo, h, u, sig, 0.3, 0.3, 0.0349065)
    ok = ~(np.isnan(wa).any(axis=1) | np.isnan(wi).any(axis=1))
    assert ok.mean() > 0.98
    # the sampled direction: 2e-4 absolute on >= 99.5 % (a lobe pick can flip on an ulp of the attenuation pdf)
    d = np.abs(wi[ok] - wa[ok]).max(axis=1)
    assert (d <= 2e-4).mean() >= 0.995, (d <= 2e-4).mean()
    same = ok.copy(); same[ok] = d <= 2e-4
    _close(f[same], fa[same], "f(sampled)")
    _close(pdf[same], pa[same], "pdf(sampled)")


def test_bsdf_hooks_reject_bad_arguments():
    z = np.zeros((4, 3), np.float32)
    with pytest.raises(api.HairMSNNError):
        api.bsdf_eval(z, z, np.zeros(4, np.float32), device=99)
