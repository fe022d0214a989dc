"""Throughput of the three entry points on the bench scene (BASELINE.json configs 2-4): Mpaths/s of
render_path_tracing, render_nrc, render_hair_msnn BETA=1 and BETA=10 at 1024x1024, one B200.
usage: python scripts/measure_modes.py [frames]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from hairmsnn_b200 import api

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 32
sc, kw = bench.make_scene(50000)
W, H = bench.W, bench.H
out = {}
for name, kind, beta in (("render_path_tracing", api.PATH_TRACING, 1), ("render_nrc", api.NRC, 1),
                         ("render_hair_msnn BETA=1", api.HAIR_MSNN, 1), ("render_hair_msnn BETA=10", api.HAIR_MSNN, 10)):
    r = api.Renderer(sc, kind, beta_cli=beta)
    r.render_frames(8)
    r.sync()
    r.reset_stats()
    r.set_profiling(True)
    r.set_collect_stats(True)
    t0 = time.perf_counter()
    r.render_frames_async(frames)
    r.sync()
    dt = time.perf_counter() - t0
    st = r.stats()
    out[name] = {"mpaths_per_s": W * H * frames / dt / 1e6, "ms_per_frame": dt / frames * 1e3,
                 "rays_per_frame": (st.rays_primary + st.rays_extend + st.rays_shadow) / max(st.frames, 1),
                 "stage_ms_per_frame": {k: round(getattr(st, "ms_" + k) / max(st.frames, 1), 3) for k in
                                        ("primary", "shade", "extend", "shadow", "finalize", "train", "infer", "composite")},
                 "launches_per_frame": st.kernel_launches / max(st.frames, 1)}
    print(name, json.dumps(out[name]), flush=True)
    r.close()
print(json.dumps(out))
