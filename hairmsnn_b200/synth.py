"""Procedural stand-ins for the reference's scene assets (the .hair / head.obj / env EXR
files are not redistributable and are absent on the benchmark box).  Same shapes and
scales as scenes/curly and scenes/straight: 50 000 strands, 3.39 M / 1.2 M
Catmull-Rom segments of radius 0.02 around a ~25-unit head in a ~190-unit scene, a
4096x2048 RGBA32F lat-long environment with a compact bright source, camera from
config.json.  Everything is generated with numpy from fixed seeds."""
import numpy as np

CAMERA_FROM = (-221.48236083984375, -1.6946277618408203, 6.213634490966797)
COS_FOVY = 0.6600000262260437


def make_hair(num_strands=50000, segs_per_strand=68, curly=True, seed=7, head_radius=24.0, length=70.0,
              thickness=0.1):
    """Returns (control_points [n,4] incl. phantom endpoints, segment_first_cp [m]) laid out
    as Scene::extractHairData does (scene.cpp:10-73): radius = 0.2 * thickness."""
    rng = np.random.default_rng(seed)
    S, K = num_strands, segs_per_strand + 1
    # roots on the upper/back part of the scalp
    z = rng.uniform(0.05, 1.0, S)
    phi = rng.uniform(0, 2 * np.pi, S)
    r = np.sqrt(1 - z * z)
    root_dir = np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1)
    root = root_dir * head_radius + np.array([0.0, 0.0, 8.0])
    t = np.linspace(0.0, 1.0, K)[None, :, None]                       # [1,K,1]
    L = length * rng.uniform(0.8, 1.1, (S, 1, 1))
    out_dir = root_dir[:, None, :]
    # grow outwards, then fall
    pos = root[:, None, :] + out_dir * (L * 0.35 * (1 - (1 - t) ** 2)) + np.array([0, 0, -1.0])[None, None, :] * (L * 0.9 * t * t)
    # frame for curls
    up = np.array([0.0, 0.0, 1.0])
    e1 = np.cross(root_dir, up); e1 /= np.linalg.norm(e1, axis=1, keepdims=True) + 1e-9
    e2 = np.cross(root_dir, e1)
    if curly:
        turns = rng.uniform(6.0, 10.0, (S, 1, 1))
        ph = rng.uniform(0, 2 * np.pi, (S, 1, 1))
        rad = rng.uniform(1.2, 2.5, (S, 1, 1)) * np.minimum(1.0, 4 * t)
        ang = 2 * np.pi * turns * t + ph
        pos = pos + e1[:, None, :] * (rad * np.cos(ang)) + e2[:, None, :] * (rad * np.sin(ang))
    else:
        wob = rng.uniform(0.0, 0.6, (S, 1, 1))
        ph = rng.uniform(0, 2 * np.pi, (S, 1, 1))
        pos = pos + e1[:, None, :] * (wob * np.sin(3 * np.pi * t + ph))
    pos = pos.astype(np.float32)
    # phantom endpoints: p0 + (p0 - p1), pN + (pN - pN-1)
    first = pos[:, :1] + (pos[:, :1] - pos[:, 1:2])
    last = pos[:, -1:] + (pos[:, -1:] - pos[:, -2:-1])
    cps = np.concatenate([first, pos, last], axis=1)                   # [S, K+2, 3]
    w = np.full(cps.shape[:2] + (1,), np.float32(0.2) * np.float32(thickness), np.float32)
    cps = np.concatenate([cps, w], axis=2).reshape(-1, 4).astype(np.float32)
    base = (np.arange(S, dtype=np.int64) * (K + 2))[:, None] + np.arange(segs_per_strand, dtype=np.int64)[None, :]
    return cps, base.reshape(-1).astype(np.int32)


def make_head(radius=24.0, n_lat=140, n_lon=280, center=(0.0, 0.0, 8.0)):
    """UV-sphere head: 2*n_lat*n_lon-ish triangles (default ~78 k like head.obj), flattened soup."""
    th = np.linspace(0, np.pi, n_lat + 1)
    ph = np.linspace(0, 2 * np.pi, n_lon + 1)
    T, P = np.meshgrid(th, ph, indexing="ij")
    n = np.stack([np.sin(T) * np.cos(P), np.sin(T) * np.sin(P), np.cos(T)], axis=2)
    v = n * radius * np.array([1.0, 0.85, 1.15]) + np.array(center)
    nn = n / np.array([1.0, 0.85, 1.15]); nn /= np.linalg.norm(nn, axis=2, keepdims=True)
    a, b, c, d = (0, 0), (1, 0), (1, 1), (0, 1)

    def corner(arr, o):
        return arr[o[0]:n_lat + o[0], o[1]:n_lon + o[1]].reshape(-1, 3)
    tv = np.stack([corner(v, a), corner(v, b), corner(v, c), corner(v, a), corner(v, c), corner(v, d)], axis=1).reshape(-1, 3)
    tn = np.stack([corner(nn, a), corner(nn, b), corner(nn, c), corner(nn, a), corner(nn, c), corner(nn, d)], axis=1).reshape(-1, 3)
    # drop degenerate pole triangles
    tri = tv.reshape(-1, 3, 3)
    area = np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1)
    keep = area > 1e-6
    return tv.reshape(-1, 3, 3)[keep].reshape(-1, 3).astype(np.float32), tn.reshape(-1, 3, 3)[keep].reshape(-1, 3).astype(np.float32)


def make_env(width=4096, height=2048, seed=11):
    """Studio-like lat-long environment: dim gradient + two soft boxes + one compact key light."""
    rng = np.random.default_rng(seed)
    v = (np.arange(height, dtype=np.float32) + 0.5) / height
    u = (np.arange(width, dtype=np.float32) + 0.5) / width
    U, V = np.meshgrid(u, v)
    base = 0.08 + 0.25 * (1 - V) ** 2
    img = np.stack([base * 1.0, base * 0.97, base * 0.92], axis=2).astype(np.float32)

    def blob(cu, cv, su, sv, rgb):
        du = np.minimum(np.abs(U - cu), 1 - np.abs(U - cu))
        g = np.exp(-0.5 * ((du / su) ** 2 + ((V - cv) / sv) ** 2)).astype(np.float32)
        return g[:, :, None] * np.array(rgb, np.float32)[None, None, :]
    img += blob(0.25, 0.35, 0.05, 0.06, (6.0, 5.6, 5.0))
    img += blob(0.70, 0.40, 0.08, 0.05, (2.5, 2.8, 3.2))
    img += blob(0.52, 0.22, 0.008, 0.01, (90.0, 85.0, 70.0))
    img *= (1.0 + 0.05 * rng.standard_normal((height // 16, width // 16, 1)).repeat(16, axis=0).repeat(16, axis=1)).astype(np.float32)
    img = np.maximum(img, 1e-3)
    return np.concatenate([img, np.ones((height, width, 1), np.float32)], axis=2).astype(np.float32)


def scene_kwargs(kind="curly", width=1024, height=1024, spp=1, num_strands=50000, env_size=(4096, 2048),
                 head=True, path_v2=40):
    """Keyword arguments for api.Scene.from_arrays mirroring scenes/<kind>/config.json."""
    segs = 68 if kind == "curly" else 24
    cps, seg = make_hair(num_strands, segs, curly=(kind == "curly"))
    kw = dict(control_points=cps, segment_first_cp=seg, num_strands=num_strands,
              cam_from=CAMERA_FROM, cam_to=(0, 0, 0), cam_up=(0, 0, 1), cos_fovy=COS_FOVY,
              sigma_a=(0.06, 0.1, 0.2), beta_m=0.3, beta_n=0.3, alpha_deg=2.0,
              env_rgba=make_env(*env_size), env_scale=1.0,
              dl_from=[(3, 3, 3)], dl_emit=[(1, 1, 1)],
              width=width, height=height, spp=spp, path_v1=1, path_v2=path_v2, mis=True, env_pdf=True)
    if head:
        tv, tn = make_head()
        kw.update(tri_vertices=tv, tri_normals=tn)
    return kw
