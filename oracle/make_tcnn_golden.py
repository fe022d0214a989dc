#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Reduces the raw dump of oracle/_ref/tcnn_golden (the reference's own tiny-cuda-nn,
run on a B200 through gpurun) to the small fixtures tests/golden/tcnn_<in_ch>.npz that pin
oracle/mlp_oracle.py and the CUDA network (tests/test_cpu_mlp_oracle.py, tests/test_gpu_mlp.py).

    python oracle/make_tcnn_golden.py gpurun_out/tcnn12 12 tests/golden/tcnn_12.npz

Per-row outputs keep the first ROWS rows, parameter-sized arrays keep the MLP matrices in full and a strided
sample of the hash grid plus float64 sums over everything.
"""
import sys

import numpy as np

ROWS = 4096
N = 16384
MLP_PARAMS = 64 * 64 + 64 * 64 + 16 * 64     # tcnn order: network first, then the encoding's grid
GRID_STRIDE = 37


def rd(d, name, dt):
    return np.fromfile(f"{d}/{name}", dtype=dt)


def reduce_params(a):
    a64 = a.astype(np.float64)
    return {"mlp": a[:MLP_PARAMS].copy(), "grid_sample": a[MLP_PARAMS::GRID_STRIDE].copy(),
            "sum": np.float64(a64.sum()), "abs_sum": np.float64(np.abs(a64).sum()), "sq_sum": np.float64((a64 * a64).sum()),
            "nonzero": np.int64(np.count_nonzero(a))}


def main():
    d, in_ch, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
    import tcnn_inputs
    assert np.array_equal(rd(d, "inputs.bin", np.float32).reshape(N, in_ch), tcnn_inputs.make_inputs(N, in_ch)), "input restatement differs"
    assert np.array_equal(rd(d, "targets.bin", np.float32).reshape(N, 3), tcnn_inputs.make_targets(N)), "target restatement differs"
    fx = {"in_ch": np.int32(in_ch), "n_rows": np.int32(N), "rows_kept": np.int32(ROWS), "grid_stride": np.int32(GRID_STRIDE),
          "input_seed": np.uint32(12345), "target_seed": np.uint32(777)}
    for tag, dt in (("params0_f32", np.float32), ("params0_f16", np.float16), ("params1_f32", np.float32),
                    ("params1_f16", np.float16), ("params4_f32", np.float32), ("grad1_f16", np.float16),
                    ("params_reset1_f32", np.float32)):
        a = rd(d, tag + ".bin", dt)
        fx["n_params"] = np.int64(a.size)
        for k, v in reduce_params(a).items():
            fx[f"{tag}.{k}"] = v
    for tag in ("infer0", "infer4"):
        a = rd(d, tag + ".bin", np.float32).reshape(N, 3)
        fx[tag] = a[:ROWS].copy()
        fx[tag + ".sum"] = a.astype(np.float64).sum(axis=0)
        fx[tag + ".sq_sum"] = (a.astype(np.float64) ** 2).sum(axis=0)
    for tag in ("fwd1_out_f16", "dLdo1_f16"):
        a = rd(d, tag + ".bin", np.float16).reshape(N, 16)      # GPUMatrix<T>(16, N), column-major
        fx[tag] = a[:ROWS, :3].copy()
        fx[tag + ".pad_abs_max"] = np.float64(np.abs(a[:, 3:].astype(np.float64)).max())
        fx[tag + ".sum"] = a[:, :3].astype(np.float64).sum(axis=0)
    fx["losses"] = rd(d, "losses.bin", np.float32)
    fx["loss_reset"] = rd(d, "loss_reset.bin", np.float32)
    np.savez_compressed(out, **fx)
    print(out, {k: (v.shape if hasattr(v, "shape") and v.shape else v) for k, v in fx.items() if "." not in k})


if __name__ == "__main__":
    main()
