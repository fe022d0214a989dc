"""Run-to-run determinism of the path tracer (PT accumulation, HairMSNN path-traced component and training records)
and of the network-dependent outputs, for the library in HM_LIB."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from hairmsnn_b200 import api
from common import small_scene_kwargs
W, H = 256, 128
kw = small_scene_kwargs(width=W, height=H, strands=1500, segs=16, path_v2=10)
sc = api.Scene.from_arrays(**kw)
tag = os.environ.get("HM_LIB", "default").split("_")[-1]
ref = None
for rep in range(6):
    pt = api.Renderer(sc, api.PATH_TRACING)
    pt.render_frames(3)
    a = pt.buffer(api.BUF_FINAL_ACCUM).copy()
    pt.close()
    r = api.Renderer(sc, api.HAIR_MSNN, beta_cli=1)
    r.render_frames(4)
    cur = (a, r.buffer(api.BUF_PT_ACCUM).copy(), r.buffer(api.BUF_NN_TRAIN_OUTPUT).copy(), r.buffer(api.BUF_FINAL_ACCUM).copy(),
           r.mlp().get_params().copy())
    r.close()
    if ref is None:
        ref = cur
        continue
    names = ("pt_renderer", "msnn_pt_accum", "train_targets", "msnn_final", "params")
    print(tag, rep, {n: (int((x != y).sum()), float(np.abs(x - y).max())) for n, x, y in zip(names, ref, cur)}, flush=True)
