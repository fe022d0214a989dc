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


def probe_scene_desc(scene, kw):
    """ctypes struct for tests/cpu_probe.cpp's ProbeSceneDesc from an api.Scene."""
    import ctypes as C
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)

    class D(C.Structure):
        _fields_ = [("nodes", fp), ("num_nodes", C.c_int), ("leaf_code", ip), ("leaf_prim", ip), ("leaf_data", fp),
                    ("cps", fp), ("tri_verts", fp), ("tri_normals", fp), ("seg_cp", ip),
                    ("num_segments", C.c_int), ("num_tris", C.c_int),
                    ("env", fp), ("cpdf", fp), ("ccdf", fp), ("mpdf", fp), ("mcdf", fp),
                    ("env_w", C.c_int), ("env_h", C.c_int), ("env_scale", C.c_float), ("env_rot", C.c_float),
                    ("has_env", C.c_int), ("env_pdf", C.c_int),
                    ("num_dlights", C.c_int), ("dl_from", fp), ("dl_emit", fp),
                    ("sigma_a", C.c_float * 3), ("beta_m", C.c_float), ("beta_n", C.c_float), ("alpha", C.c_float),
                    ("gains", C.c_float * 4), ("kd", C.c_float * 3), ("surf_alpha", C.c_float),
                    ("scene_scale", C.c_float), ("mis", C.c_int),
                    ("cam_pos", C.c_float * 3), ("cam_d00", C.c_float * 3), ("cam_du", C.c_float * 3), ("cam_dv", C.c_float * 3)]
    info = scene.info()
    arr = scene.arrays()
    d = D()
    keep = [arr]
    P = lambda a, t=fp: a.ctypes.data_as(t)
    d.nodes = P(arr["nodes"]); d.num_nodes = info.num_bvh_nodes
    d.leaf_code = P(arr["leaf_code"], ip); d.leaf_prim = P(arr["leaf_prim"], ip); d.leaf_data = P(arr["leaf_data"])
    d.cps = P(arr["cps"]); d.tri_verts = P(arr["tri_verts"]); d.tri_normals = P(arr["tri_normals"]); d.seg_cp = P(arr["seg_cp"], ip)
    d.num_segments = info.num_segments; d.num_tris = info.num_triangles
    d.has_env = int(info.env_w > 0)
    if d.has_env:
        t = scene.env_tables(); keep.append(t)
        d.env = P(t["env"]); d.cpdf = P(t["cpdf"]); d.ccdf = P(t["ccdf"]); d.mpdf = P(t["mpdf"]); d.mcdf = P(t["mcdf"])
        d.env_w, d.env_h = info.env_w, info.env_h
    d.env_scale = kw.get("env_scale", 1.0); d.env_rot = kw.get("env_rotation", 0.0); d.env_pdf = int(kw.get("env_pdf", True))
    dlf = np.ascontiguousarray(kw.get("dl_from", ()), np.float32).reshape(-1, 3)
    if len(dlf):
        dlf = (dlf / np.sqrt((dlf ** 2).sum(axis=1, keepdims=True, dtype=np.float32))).astype(np.float32)
    dle = np.ascontiguousarray(kw.get("dl_emit", ()), np.float32).reshape(-1, 3)
    keep += [dlf, dle]
    d.num_dlights = len(dlf)
    if len(dlf):
        d.dl_from = P(dlf); d.dl_emit = P(dle)
    d.sigma_a = (C.c_float * 3)(*kw.get("sigma_a", (0.06, 0.1, 0.2)))
    d.beta_m = kw.get("beta_m", 0.3); d.beta_n = kw.get("beta_n", 0.3)
    d.alpha = float(np.float32(3.14159) * np.float32(kw.get("alpha_deg", 2.0)) / np.float32(180.0))
    d.gains = (C.c_float * 4)(*kw.get("gains", (1, 1, 1, 1)))
    d.kd = (C.c_float * 3)(*kw.get("surface_kd", (0, 0, 0))); d.surf_alpha = kw.get("surface_alpha", 1.0)
    d.scene_scale = info.scene_scale; d.mis = int(kw.get("mis", True))
    d.cam_pos = info.cam_pos; d.cam_d00 = info.cam_d00; d.cam_du = info.cam_du; d.cam_dv = info.cam_dv
    d._keep = keep
    return d
