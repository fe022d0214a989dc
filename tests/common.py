"""Shared helpers for the test-suite: small deterministic scenes."""
import numpy as np

from hairmsnn_b200 import synth


def small_scene_kwargs(width=128, height=128, strands=600, segs=12, env=(256, 128), head=True, curly=True, path_v2=40,
                       mis=True, env_pdf=True, dlights=True, env_light=True, seed=3, thickness=2.0):
    cps, seg = synth.make_hair(strands, segs, curly=curly, seed=seed, thickness=thickness)
    kw = dict(control_points=cps, segment_first_cp=seg, num_strands=strands,
              cam_from=synth.CAMERA_FROM, cam_to=(0, 0, 0), cam_up=(0, 0, 1), cos_fovy=synth.COS_FOVY,
              sigma_a=(0.06, 0.1, 0.2), beta_m=0.3, beta_n=0.3, alpha_deg=2.0,
              width=width, height=height, spp=1, path_v1=1, path_v2=path_v2, mis=mis, env_pdf=env_pdf)
    if env_light:
        kw.update(env_rgba=synth.make_env(*env), env_scale=1.0)
    if dlights:
        kw.update(dl_from=[(3, 3, 3)], dl_emit=[(1, 1, 1)])
    if head:
        tv, tn = synth.make_head(n_lat=24, n_lon=48)
        kw.update(tri_vertices=tv, tri_normals=tn)
    return kw


def camera_rays(info, n, seed=0):
    """n primary-like rays through random screen positions."""
    rng = np.random.default_rng(seed)
    su = rng.random(n).astype(np.float32); sv = rng.random(n).astype(np.float32)
    d00 = np.array(info.cam_d00[:], np.float32); du = np.array(info.cam_du[:], np.float32); dv = np.array(info.cam_dv[:], np.float32)
    d = d00[None] + su[:, None] * du[None] + sv[:, None] * dv[None]
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    o = np.tile(np.array(info.cam_pos[:], np.float32), (n, 1))
    return o, d
