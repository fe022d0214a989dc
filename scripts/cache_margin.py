"""Margin of tests/test_gpu_msnn.py::test_cache_learns_the_residual (e_final / e_short), printed for the library in HM_LIB."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from hairmsnn_b200 import api
from common import small_scene_kwargs
W, H = 256, 128
kw = small_scene_kwargs(width=W, height=H, strands=1500, segs=16, path_v2=10)
sc = api.Scene.from_arrays(**kw)
pt = api.Renderer(sc, api.PATH_TRACING)
pt.render_frames(48)
truth = pt.buffer(api.BUF_FINAL_AVG)[..., :3]
for rep in range(3):
    r = api.Renderer(sc, api.HAIR_MSNN, beta_cli=1)
    r.msnn_pretrain(300)
    r.render_frames(48)
    final = r.buffer(api.BUF_FINAL_AVG)[..., :3]
    short = r.buffer(api.BUF_PT_AVG)[..., :3]
    gb = r.buffer(api.BUF_GBUFFER).reshape(H, W, 4)
    flags = gb[..., 3].copy().view(np.int32)
    hair = ((flags & 1) != 0) & ((flags & 2) == 0)
    e_short = np.abs(short[hair].mean(axis=0) - truth[hair].mean(axis=0)).sum()
    e_final = np.abs(final[hair].mean(axis=0) - truth[hair].mean(axis=0)).sum()
    print(os.environ.get("HM_LIB", "default").split("_")[-1], "e_short", e_short, "e_final", e_final, "ratio", e_final / e_short, flush=True)
    r.close()
