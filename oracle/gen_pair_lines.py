#!/usr/bin/env python3
"""TEST INFRASTRUCTURE (oracle).  Scan the reference's per-path sources for source
lines holding TWO lcg_randomf() calls (constructor argument lists such as
vec2f(lcg_randomf(rng), lcg_randomf(rng)), optix_common.cuh:91).  g++ evaluates those
arguments right-to-left while the shipped GPU binary (nvcc device code) evaluates
them left-to-right; oracle/ref_optix_emul reorders the two draws on
exactly these lines so the host build reproduces the GPU streams.

usage: gen_pair_lines.py <reference root> <out.inc>
"""
import os
import re
import sys

FILES = ["cuda_headers/optix_common.cuh", "cuda/path_tracing.cu", "cuda/hair_msnn.cu", "cuda/nrc.cu"]


def main():
    root, out = sys.argv[1], sys.argv[2]
    rows = []
    for rel in FILES:
        with open(os.path.join(root, rel)) as f:
            for ln, line in enumerate(f, 1):
                n = len(re.findall(r"lcg_randomf\s*\(", line))
                if "__device__" in line:
                    continue
                if n == 2:
                    rows.append((os.path.basename(rel), ln))
                elif n > 2:
                    raise SystemExit(f"{rel}:{ln}: {n} draws on one line - extend the reordering shim")
    with open(out, "w") as f:
        for name, ln in rows:
            f.write('{"%s", %d},\n' % (name, ln))
    print(f"{len(rows)} paired-draw lines -> {out}")


if __name__ == "__main__":
    main()
