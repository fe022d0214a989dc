"""A/B of the tail-group size (HM_TAIL_GROUP: tail pieces of that many consecutive HairMSNN frames run as one launch
sequence) and the frames in flight, on the bench workload, one scene load for all settings.
usage: python scripts/tail_group_sweep.py [steps] "G:F,G:F,..."   (F = 0: the library's default for that G)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from hairmsnn_b200 import api

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 24
settings = sys.argv[2] if len(sys.argv) > 2 else "1:8,4:12,1:8,4:12"
torch.cuda.set_device(0)
peaks = bench.load_peaks()
sc, kw, data, W, H = bench.make_scene("msnn_b1", 50000)
for item in settings.split(","):
    g, f = (int(x) for x in item.split(":"))
    os.environ["HM_TAIL_GROUP"] = str(g)
    if f:
        os.environ["HM_FRAMES_IN_FLIGHT"] = str(f)
    else:
        os.environ.pop("HM_FRAMES_IN_FLIGHT", None)
    r = api.Renderer(sc, api.HAIR_MSNN, beta_cli=1)
    m = bench.measure(r, api, torch, 0, None, W, H, steps, 6, peaks, "msnn")
    st = m["stage_ms_per_step"]
    print(f"group {g} in_flight {f or 'default'}: value {m['value']:.1f} Mpaths/s  ms/step {m['ms'] / steps:.3f}  e2e {m['e2e']['value']:.1f}  "
          f"launches/step {m['launches'] / steps:.0f}  tail {st['tail_piece']:.1f} trace {st['trace']:.2f} train {st['train']:.2f} infer {st['infer']:.2f}  loss {m['loss']:.3f}",
          flush=True)
    r.close()
