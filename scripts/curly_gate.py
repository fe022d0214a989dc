#!/usr/bin/env python
"""Image gate of the headline config.

Loads a shipped scene through hm_scene_load (assets/scenes/<scene>/config.json, staged by
scripts/stage_assets.py), renders
  * render_path_tracing, <pt_spp> samples          -> the ground truth (config 2),
  * render_path_tracing, another <spp> samples with disjoint sample ids -> the noise floor of relMSE at that spp,
  * render_hair_msnn BETA in --betas, <spp> samples after the initial training (config 4),
and reports relMSE = mean over pixels and RGB of (I - R)^2 / (R^2 + 0.01) of each against the ground truth,
plus Mpaths/s of each render.  Writes <out>/gate.json, PNGs and (optionally) EXRs.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


def rel_mse(img, ref, eps=1e-2):
    img = np.asarray(img, np.float64)[..., :3]
    ref = np.asarray(ref, np.float64)[..., :3]
    return float(np.mean((img - ref) ** 2 / (ref ** 2 + eps)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="curly")
    ap.add_argument("--config", default=None)
    ap.add_argument("--pt-spp", type=int, default=500)
    ap.add_argument("--spp", type=int, default=500)
    ap.add_argument("--betas", default="1,10")
    ap.add_argument("--pretrain", type=int, default=200)
    ap.add_argument("--nrc", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "curly_gate"))
    ap.add_argument("--exr", action="store_true")
    args = ap.parse_args()
    from hairmsnn_b200 import api

    cfg = args.config or os.path.join(ROOT, "assets", "scenes", args.scene, "config.json")
    os.makedirs(args.out, exist_ok=True)
    t0 = time.time()
    sc = api.Scene.load(cfg)
    info = sc.info()
    W, H = info.width, info.height
    res = {"scene": cfg, "width": W, "height": H, "segments": info.num_segments, "triangles": info.num_triangles,
           "wide_nodes": info.num_wide_nodes, "leaf_refs": info.num_wide_leaf_refs, "load_s": time.time() - t0,
           "relmse_def": "mean over pixels and RGB of (I-R)^2/(R^2+0.01), R = render_path_tracing at pt_spp",
           "pt_spp": args.pt_spp, "spp": args.spp, "renders": {}}
    print(f"scene loaded in {res['load_s']:.1f}s: {info.num_segments} segments", flush=True)

    def run(name, kind, beta=1, spp=1, offset=0, pretrain=0):
        r = api.Renderer(sc, kind, beta_cli=beta, device=0)
        if offset:
            r.set_frame_schedule(offset, 1)
        tp = 0.0
        if pretrain and kind == api.HAIR_MSNN:
            t = time.perf_counter()
            r.msnn_pretrain(pretrain)
            tp = time.perf_counter() - t
        r.sync()
        t = time.perf_counter()
        done = 0
        while done < spp:
            n = min(32, spp - done)
            r.render_frames_async(n)
            done += n
        r.sync()
        dt = time.perf_counter() - t
        img = r.buffer(api.BUF_FINAL_AVG)
        st = r.stats()
        r.save_png(os.path.join(args.out, f"{name}.png"))
        if args.exr:
            r.save_exr(os.path.join(args.out, f"{name}.exr"))
        out = {"spp": spp, "seconds": dt, "mpaths_per_s": W * H * spp / dt / 1e6, "pretrain_s": tp, "pretrain_steps": pretrain,
               "mean": [float(x) for x in img[..., :3].reshape(-1, 3).mean(axis=0)], "loss": float(st.last_loss),
               "finite": bool(np.isfinite(img).all())}
        extra = None
        if kind == api.HAIR_MSNN:
            extra = (r.buffer(api.BUF_PT_AVG), r.buffer(api.BUF_NN_AVG))
        r.close()
        print(f"{name}: {spp} spp in {dt:.2f}s = {out['mpaths_per_s']:.1f} Mpaths/s", flush=True)
        return img, out, extra

    gt, o, _ = run("pt_gt", api.PATH_TRACING, spp=args.pt_spp)
    res["renders"]["pt_gt"] = o
    np.save(os.path.join(args.out, "pt_gt_small.npy"), gt[::8, ::8, :3].astype(np.float16))
    img, o, _ = run("pt_other_samples", api.PATH_TRACING, spp=args.spp, offset=100000)
    o["relmse"] = rel_mse(img, gt)
    res["renders"]["pt_other_samples"] = o
    for b in [int(x) for x in args.betas.split(",") if x]:
        img, o, extra = run(f"msnn_beta{b}", api.HAIR_MSNN, beta=b, spp=args.spp, pretrain=args.pretrain)
        o["relmse"] = rel_mse(img, gt)
        o["relmse_pt_part_only"] = rel_mse(extra[0], gt)
        res["renders"][f"msnn_beta{b}"] = o
        print(f"  relMSE beta={b}: {o['relmse']:.5f} (pt part alone {o['relmse_pt_part_only']:.5f})", flush=True)
    if args.nrc:
        img, o, _ = run("nrc", api.NRC, spp=args.spp)
        o["relmse"] = rel_mse(img, gt)
        res["renders"]["nrc"] = o
    json.dump(res, open(os.path.join(args.out, "gate.json"), "w"), indent=1)
    print(json.dumps(res["renders"], indent=1))


if __name__ == "__main__":
    main()
