"""Scratch profiling driver: full-size synthetic scene, per-stage timings."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hairmsnn_b200 import api, synth

kind = sys.argv[1] if len(sys.argv) > 1 else "curly"
mode = sys.argv[2] if len(sys.argv) > 2 else "pt"
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 4
strands = int(sys.argv[4]) if len(sys.argv) > 4 else 50000
t0 = time.time()
kw = synth.scene_kwargs(kind, 1024, 1024, num_strands=strands)
t1 = time.time()
sc = api.Scene.from_arrays(**kw)
t2 = time.time()
i = sc.info()
print(f"scene: {i.num_segments} segs {i.num_triangles} tris {i.num_bvh_nodes} nodes scale {i.scene_scale:.1f}; gen {t1-t0:.1f}s build {t2-t1:.1f}s", flush=True)
r = api.Renderer(sc, api.PATH_TRACING if mode == "pt" else api.HAIR_MSNN, beta_cli=1)
r.render_frames(1)
r.set_profiling(True)
t = time.time()
r.render_frames(frames)
dt = time.time() - t
s = r.stats()
print(f"{frames} frames {dt*1e3/frames:.2f} ms/frame wall -> {1024*1024*frames/dt/1e6:.2f} Mpaths/s (profiling on)")
for k in ("ms_primary", "ms_shade", "ms_extend", "ms_shadow", "ms_finalize", "ms_train", "ms_infer", "ms_composite", "ms_total"):
    print(f"  {k}: {getattr(s, k)/frames:.3f} ms/frame")
print(f"  rays/frame: primary {s.rays_primary/ (frames+1):.0f} extend {s.rays_extend/(frames):.0f} shadow {s.rays_shadow/(frames):.0f} shade {s.shade_items/frames:.0f}")
r.set_profiling(False)
r.sync()
t = time.time()
r.render_frames(frames)
dt = time.time() - t
print(f"no-profiling: {dt*1e3/frames:.2f} ms/frame -> {1024*1024*frames/dt/1e6:.2f} Mpaths/s")
img = r.buffer(api.BUF_FINAL_AVG)
print("mean rgb", img[..., :3].mean(axis=(0, 1)), "hit frac", float((img[..., :3].sum(axis=2) > 0).mean()))
os.makedirs("gpurun_out", exist_ok=True)
r.save_png("gpurun_out/quick_%s_%s.png" % (kind, mode))
# primary-ray traversal stats
o = np.tile(np.array(i.cam_pos[:], np.float32), (65536, 1))
rng = np.random.default_rng(0)
su, sv = rng.random(65536).astype(np.float32), rng.random(65536).astype(np.float32)
d = np.array(i.cam_d00[:], np.float32)[None] + su[:, None] * np.array(i.cam_du[:], np.float32)[None] + sv[:, None] * np.array(i.cam_dv[:], np.float32)[None]
d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
h = r.trace_rays(o, d, stats=True)
print("primary stats: hit", float((h["prim"] >= 0).mean()), "nodes/ray", h["nodes"].mean(), "prims/ray", h["prims"].mean())
