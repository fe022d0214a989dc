import sys, numpy as np
sys.path.insert(0, "tests")
from hairmsnn_b200 import api
from common import small_scene_kwargs
W, H = 256, 128
kw = small_scene_kwargs(width=W, height=H, strands=1500, segs=16, path_v2=10)
sc = api.Scene.from_arrays(**kw)
pt = api.Renderer(sc, api.PATH_TRACING); pt.render_frames(128)
truth = pt.buffer(api.BUF_FINAL_AVG)[..., :3]
kw40 = dict(kw); kw40["path_v2"] = 40
sc40 = api.Scene.from_arrays(**kw40)
pt40 = api.Renderer(sc40, api.PATH_TRACING); pt40.render_frames(128)
truth40 = pt40.buffer(api.BUF_FINAL_AVG)[..., :3]
import os
r = api.Renderer(sc, api.NRC)
r.nrc_set_all_unbiased(bool(int(os.environ.get("UNB", "0"))))
tot = 0
for n in (150, 600, 3000):
    r.render_frames(n); tot += n
    r.reset_accumulation(); r.render_frames(64); tot += 64
    img = r.buffer(api.BUF_FINAL_AVG)[..., :3]
    gb = r.buffer(api.BUF_GBUFFER).reshape(H, W, 4)
    gbb = r.buffer(api.BUF_GBUFFER_B).reshape(H, W, 4)
    hit = (gb[..., 3].copy().view(np.int32) & 1) != 0
    print(tot, "nrc", img[hit].mean(), "pt(v2=10)", truth[hit].mean(), "pt(v2=40)", truth40[hit].mean(), "short-only(last frame)", gb[hit][:, :3].mean(),
          "mean bounces", gbb[hit][:, 3].copy().view(np.int32).mean(), "loss", r.stats().last_loss)
