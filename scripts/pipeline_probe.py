"""Where does the frame time go?  Frame loop with / without the tail piece and the training step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hairmsnn_b200 import api, synth
import bench
W, H = bench.W, bench.H
for label, v2 in (("full (path_v2=40)", 40), ("no tail piece (path_v2=2)", 2)):
    kw = synth.scene_kwargs("curly", W, H, num_strands=50000)
    kw["path_v2"] = v2
    sc = api.Scene.from_arrays(**kw)
    r = api.Renderer(sc, api.HAIR_MSNN, beta_cli=1)
    r.render_frames(8); r.sync()
    t0 = time.perf_counter(); r.render_frames_async(32); r.sync(); dt = (time.perf_counter() - t0) / 32 * 1e3
    print(f"{label}: {dt:.3f} ms/frame", flush=True)
    # split calls: trace only (no training, no inference)
    t0 = time.perf_counter()
    for _ in range(32):
        r.msnn_trace(); r.msnn_finish()
    r.sync(); dt = (time.perf_counter() - t0) / 32 * 1e3
    print(f"{label}, no training step: {dt:.3f} ms/frame", flush=True)
    r.close()
