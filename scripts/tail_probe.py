"""Where does the frame time go on the real scene?  render_hair_msnn BETA=1 on scenes/curly with the full path length
(path_v2 = 40: the 16384 training paths run on in the tail piece) and with path_v2 = 2 (no tail piece at all)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("HM_BVH_CACHE", "/dev/shm/hm_bvh_probe")
os.makedirs(os.environ["HM_BVH_CACHE"], exist_ok=True)
from hairmsnn_b200 import api
base = os.path.join(ROOT, "assets", "scenes", "curly", "config.json")
for label, v2 in (("full (path_v2=40)", 40), ("no tail piece (path_v2=2)", 2)):
    cfg = json.load(open(base)); cfg["integrator"]["path_v2"] = v2
    p = os.path.join(os.path.dirname(base), f"config_probe_{v2}.json"); json.dump(cfg, open(p, "w"))
    sc = api.Scene.load(p)
    r = api.Renderer(sc, api.HAIR_MSNN, beta_cli=1)
    r.render_frames(8); r.sync()
    t0 = time.perf_counter(); r.render_frames_async(48); r.sync(); dt = (time.perf_counter() - t0) / 48 * 1e3
    print(f"{label}: {dt:.3f} ms/frame", flush=True)
    t0 = time.perf_counter()
    for _ in range(48):
        r.msnn_trace(); r.msnn_finish()
    r.sync(); dt = (time.perf_counter() - t0) / 48 * 1e3
    print(f"{label}, no training step: {dt:.3f} ms/frame", flush=True)
    r.close()
