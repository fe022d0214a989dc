"""Stand-alone timing of the radiance-cache network on the synthetic inputs of SURVEY §8d: inference of
2^20 rows (AoS fp32 [N][12]) and a 16384-row training step; CUDA events on the network's stream."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hairmsnn_b200 import api

N = 1 << 20
rng = np.random.default_rng(0)
x = np.zeros((N, 12), np.float32)
x[:, :3] = rng.uniform(-0.5, 0.5, (N, 3))
d = rng.normal(size=(N, 6)).astype(np.float32)
d[:, :3] /= np.linalg.norm(d[:, :3], axis=1, keepdims=True); d[:, 3:] /= np.linalg.norm(d[:, 3:], axis=1, keepdims=True)
x[:, 3:9] = d
m = api.Mlp.create()
xi = torch.from_numpy(x).cuda(); yo = torch.empty((N, 3), device="cuda")
tx = xi[:16384].contiguous(); ty = torch.rand((16384, 3), device="cuda")
st = torch.cuda.ExternalStream(m.stream)
torch.cuda.synchronize()
def timed(fn, reps):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps): fn()
    e1.record(st); e1.synchronize()
    return e0.elapsed_time(e1) / reps
ms = timed(lambda: m.inference_device(xi.data_ptr(), yo.data_ptr(), N), 20)
print(f"inference: {ms:.3f} ms / 2^20 queries = {N / ms / 1e6:.2f} G queries/s, {N * 16768 / ms / 1e9:.1f} TFLOP/s  (HM_MLP_CTAS={os.environ.get('HM_MLP_CTAS', 'default')})")
for nrec in (16384, 65536):
    tx = xi[:nrec].contiguous(); ty = torch.rand((nrec, 3), device="cuda")
    ms = timed(lambda: m.train_step_device(tx.data_ptr(), ty.data_ptr(), nrec), 50)
    ms_fb = timed(lambda: m.forward_backward_device(tx.data_ptr(), ty.data_ptr(), nrec), 50)
    print(f"training step ({nrec} records): {ms:.3f} ms (forward+backward alone {ms_fb:.3f} ms = {nrec * 50304 / ms_fb / 1e9:.1f} TFLOP/s algorithmic)  HM_MLP_TRAIN={os.environ.get('HM_MLP_TRAIN', 'fused')}")
# pixel-coherent inputs: positions along a smooth surface, as the frame's G-buffer provides them
xs = x.copy()
g = np.arange(N)
xs[:, 0] = ((g % 1024) / 1024.0 - 0.5) * 0.8; xs[:, 1] = ((g // 1024) / 1024.0 - 0.5) * 0.8; xs[:, 2] = 0.1 * np.sin(xs[:, 0] * 9) * np.cos(xs[:, 1] * 7)
xc = torch.from_numpy(xs).cuda()
ms = timed(lambda: m.inference_device(xc.data_ptr(), yo.data_ptr(), N), 20)
print(f"inference, pixel-coherent positions: {ms:.3f} ms / 2^20 queries = {N / ms / 1e6:.2f} G queries/s")
