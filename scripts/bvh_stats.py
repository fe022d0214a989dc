"""CPU-side BVH quality probe: nodes visited / primitives tested per ray for camera rays and
for incoherent secondary rays leaving hair hit points (host build of the product traversal)."""
import ctypes as C
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from hairmsnn_b200 import synth

fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)


def build_probe():
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libcpu_probe.so")
    srcs = [os.path.join(ROOT, "tests", "cpu_probe.cpp"), os.path.join(ROOT, "hairmsnn_b200", "csrc", "hm_bvh_build.cpp")]
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-mfma", "-pthread", "-shared",
                           "-I/usr/local/cuda/include"] + srcs + ["-o", so])
    return C.CDLL(so)


def load_hair(path):
    """Cem Yuksel .hair -> control points with phantom endpoints (Scene::extractHairData)."""
    with open(path, "rb") as f:
        hdr = f.read(128)
        assert hdr[:4] == b"HAIR"
        ns, npnt, flags, dseg = np.frombuffer(hdr[4:20], np.uint32)
        dthick = np.frombuffer(hdr[20:24], np.float32)[0]
        segs = np.frombuffer(f.read(2 * ns), np.uint16).astype(np.int64) if flags & 1 else np.full(ns, dseg, np.int64)
        pts = np.frombuffer(f.read(12 * npnt), np.float32).reshape(-1, 3)
    cps = []; seg_first = []
    start = np.concatenate([[0], np.cumsum(segs + 1)])
    out_ofs = 0
    for s in range(ns):
        p = pts[start[s]:start[s + 1]]
        c = np.concatenate([p[:1] + (p[:1] - p[1:2]), p, p[-1:] + (p[-1:] - p[-2:-1])])
        cps.append(c)
        seg_first.append(out_ofs + np.arange(len(p) - 1))
        out_ofs += len(c)
    cps = np.concatenate(cps).astype(np.float32)
    w = np.full((len(cps), 1), np.float32(0.2) * np.float32(dthick), np.float32)
    return np.concatenate([cps, w], axis=1), np.concatenate(seg_first).astype(np.int32)


WIDE = os.environ.get("HM_STATS_WIDE", "1") != "0"


def trace(probe, h, o, d, any_hit=False):
    n = len(o)
    o = np.ascontiguousarray(o, np.float32); d = np.ascontiguousarray(d, np.float32)
    t = np.zeros(n, np.float32); p = np.zeros(n, np.int32); u = np.zeros(n, np.float32); v = np.zeros(n, np.float32)
    nodes = np.zeros(n, np.int32); prims = np.zeros(n, np.int32)
    (probe.probe_trace_wide if WIDE else probe.probe_trace)(h, n, o.ctypes.data_as(fp), d.ctypes.data_as(fp), C.c_float(0), C.c_float(1e30), int(any_hit),
                      t.ctypes.data_as(fp), p.ctypes.data_as(ip), u.ctypes.data_as(fp), v.ctypes.data_as(fp),
                      nodes.ctypes.data_as(ip), prims.ctypes.data_as(ip))
    return t, p, nodes, prims


def main():
    real = len(sys.argv) > 1 and sys.argv[1] == "real"
    n_rays = 20000
    probe = build_probe()
    if real:
        cps, seg = load_hair("/root/reference/scenes/curly/wCurly.hair")
    else:
        cps, seg = synth.make_hair(50000, 68, curly=True)
    tv, _ = synth.make_head()
    tv4 = np.concatenate([tv, np.zeros((len(tv), 1), np.float32)], axis=1).astype(np.float32)
    probe.probe_scene_create.restype = C.c_void_p
    t0 = time.time()
    h = C.c_void_p(probe.probe_scene_create(cps.ctypes.data_as(fp), len(cps), seg.ctypes.data_as(ip), len(seg),
                                           tv4.ctypes.data_as(fp), len(tv4) // 3, 0))
    probe.probe_scene_num_nodes.argtypes = [C.c_void_p]
    probe.probe_scene_num_wide_nodes.argtypes = [C.c_void_p]; probe.probe_scene_wide_depth.argtypes = [C.c_void_p]
    nb, nw = probe.probe_scene_num_nodes(h), probe.probe_scene_num_wide_nodes(h)
    print(f"segments {len(seg)} binary nodes {nb} ({nb * 64 / 1e6:.0f} MB) = references - 1; wide nodes {nw} ({nw * 80 / 1e6:.0f} MB), "
          f"wide leaf copies {(nb + 1) * 64 / 1e6:.0f} MB, wide depth {probe.probe_scene_wide_depth(h)}, build {time.time() - t0:.1f}s "
          f"split={os.environ.get('HM_BVH_SPLIT', 'default')} span={os.environ.get('HM_BVH_SPAN', 'default')} tree={'wide' if WIDE else 'binary'}")
    # camera rays (config.json camera, 1024x1024)
    rng = np.random.default_rng(0)
    cam = np.array(synth.CAMERA_FROM, np.float32)
    fwd = -cam / np.linalg.norm(cam)
    right = np.cross(fwd, [0, 0, 1]); right /= np.linalg.norm(right)
    upv = np.cross(right, fwd)
    su = rng.random(n_rays) - 0.5; sv = rng.random(n_rays) - 0.5
    d = fwd[None] + 0.66 * su[:, None] * right[None] + 0.66 * sv[:, None] * upv[None]
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    o = np.tile(cam, (n_rays, 1)).astype(np.float32)
    t0 = time.time()
    t, p, nodes, prims = trace(probe, h, o, d)
    print(f"primary : hit {np.mean(p >= 0):.3f} nodes/ray {nodes.mean():7.1f} prims/ray {prims.mean():6.1f} max nodes {nodes.max()}  ({time.time() - t0:.1f}s)")
    hit = (p >= 0) & (p < len(seg))
    ph = o[hit] + t[hit, None] * d[hit]
    z = rng.uniform(-1, 1, len(ph)); phi = rng.uniform(0, 2 * np.pi, len(ph)); r = np.sqrt(1 - z * z)
    d2 = np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1).astype(np.float32)
    o2 = (ph + 0.05 * d2).astype(np.float32)
    if os.environ.get("HM_STATS_SORTED"):
        for mode, what in ((0, "octant order, no culling"), (2, "octant order + pop-time culling"), (8, "octant order + group-min culling"), (4, "nearest first + octant, no culling"), (6, "nearest first + octant + culling"), (28, "nearest first + cull rest-of-group min"), (3, "distance order + pop-time culling")):
          for label, oo, dd in (("primary", o, d), ("closest", o2, d2)):
            nn = len(oo); oo = np.ascontiguousarray(oo, np.float32); dd = np.ascontiguousarray(dd, np.float32)
            pp = np.zeros(nn, np.int32); nodes = np.zeros(nn, np.int32); prims = np.zeros(nn, np.int32)
            probe.probe_trace_wide_sorted(h, nn, oo.ctypes.data_as(fp), dd.ctypes.data_as(fp), C.c_float(0), C.c_float(1e30),
                                          pp.ctypes.data_as(ip), nodes.ctypes.data_as(ip), prims.ctypes.data_as(ip), mode)
            print(f"{label:8s} {what:36s}: nodes/ray {nodes.mean():7.1f} prims/ray {prims.mean():6.1f}")
    for any_hit in (False, True):
        t0 = time.time()
        t, p, nodes, prims = trace(probe, h, o2, d2, any_hit)
        print(f"{'any-hit ' if any_hit else 'closest '}: hit {np.mean(p >= 0):.3f} nodes/ray {nodes.mean():7.1f} prims/ray {prims.mean():6.1f} max nodes {nodes.max()}  ({time.time() - t0:.1f}s)")


if __name__ == "__main__":
    main()
