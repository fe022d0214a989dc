#!/usr/bin/env python
"""Stages the reference's shipped scenes (scenes/curly, scenes/straight, scenes/envmaps: data fixtures,
not source) from /root/reference into assets/scenes/ — git-ignored, but shipped to the GPU box with the
gpurun snapshot like oracle/_ref.  Nothing on the GPU box reads /root/reference.

    python scripts/stage_assets.py [--ref /root/reference] [--force]
"""
import argparse
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "assets", "scenes")


def stage(ref="/root/reference", force=False):
    src_root = os.path.join(ref, "scenes")
    if not os.path.isdir(src_root):
        return False
    for sub in ("curly", "straight", "envmaps"):
        src, dst = os.path.join(src_root, sub), os.path.join(DEST, sub)
        os.makedirs(dst, exist_ok=True)
        for name in sorted(os.listdir(src)):
            s, d = os.path.join(src, name), os.path.join(dst, name)
            if os.path.isfile(s) and (force or not os.path.exists(d) or os.path.getsize(d) != os.path.getsize(s)):
                shutil.copyfile(s, d)
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    ok = stage(a.ref, a.force)
    print("staged into", DEST if ok else "(reference tree not found)")
    sys.exit(0 if ok else 1)
