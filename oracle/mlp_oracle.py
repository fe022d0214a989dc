"""TEST INFRASTRUCTURE — CPU restatement (numpy) of the tiny-cuda-nn network the reference
instantiates for render_hair_msnn / render_nrc.  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module;
the product never does.

PINNED against the reference's own tiny-cuda-nn: oracle/Makefile.tcnn compiles the vendored
library for sm_100, oracle/tcnn_golden.cu drives it the way TINY_MLP does on seeded inputs, and the
outputs of one run on a B200 are committed as tests/golden/tcnn_{12,9}.npz
(tests/test_cpu_mlp_oracle.py checks every function below against them).  Two things only the
golden vectors revealed: the grid initialisation is an FMA (nvcc contracts val * scale + lower), and
with 9 inputs (render_nrc) OneBlob's SoA padding overwrites network inputs 38-45 with 1.0 and
leaves 56-63 unwritten (0 in a fresh allocation).  The PRNG is also pinned against the published
PCG32 reference stream and std::seed_seq (C++ standard [rand.util.seedseq]).

All paths are relative to /root/reference/extern/tiny-cuda-nn.
Two arithmetic modes:
  half=True  : round where tcnn holds __half (parameters, encoded features incl. the fp16 accumulation of
               the 8 grid corners (grid.h:337-341), activations between layers, network output, loss
               gradients) — what tcnn computes, up to its fp16 ACCUMULATION inside wmma;
  half=False : everything in fp32 — the mathematical ground truth.
"""
import numpy as np

MASK64 = (1 << 64) - 1
PCG32_MULT = 0x5851F42D4C957F2D


class Pcg32:
    """dependencies/pcg32/pcg32.h"""

    def __init__(self, initstate=None, initseq=1):
        if initstate is None:
            self.state, self.inc = 0x853C49E6748FEA9B, 0xDA3E39CB94B95BDB
        else:
            self.state = 0
            self.inc = ((initseq << 1) | 1) & MASK64
            self.next_uint()
            self.state = (self.state + initstate) & MASK64
            self.next_uint()

    def next_uint(self):
        old = self.state
        self.state = (old * PCG32_MULT + self.inc) & MASK64
        xorshifted = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
        rot = old >> 59
        return ((xorshifted >> rot) | (xorshifted << ((-rot) & 31))) & 0xFFFFFFFF

    def next_float(self):
        u = np.array([(self.next_uint() >> 9) | 0x3F800000], np.uint32)
        return np.float32(u.view(np.float32)[0] - np.float32(1.0))

    def advance(self, delta):
        cur_mult, cur_plus, acc_mult, acc_plus = PCG32_MULT, self.inc, 1, 0
        delta &= MASK64
        while delta > 0:
            if delta & 1:
                acc_mult = (acc_mult * cur_mult) & MASK64
                acc_plus = (acc_plus * cur_mult + cur_plus) & MASK64
            cur_plus = ((cur_mult + 1) * cur_plus) & MASK64
            cur_mult = (cur_mult * cur_mult) & MASK64
            delta >>= 1
        self.state = (acc_mult * self.state + acc_plus) & MASK64

    def floats(self, n):
        """n consecutive next_float() values, vectorised."""
        out = np.empty(n, np.uint32)
        s, inc = self.state, self.inc
        for i in range(n):
            old = s
            s = (old * PCG32_MULT + inc) & MASK64
            xs = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
            rot = old >> 59
            out[i] = ((xs >> rot) | (xs << ((-rot) & 31))) & 0xFFFFFFFF
        self.state = s
        return ((out >> 9) | 0x3F800000).view(np.float32) - np.float32(1.0)


def seed_seq_generate(values, n):
    """std::seed_seq{values...}.generate(n outputs) — C++ standard [rand.util.seedseq]."""
    M = 0xFFFFFFFF
    v = [x & M for x in values]
    s = len(v)
    b = [0x8B8B8B8B] * n
    t = 11 if n >= 623 else 7 if n >= 68 else 5 if n >= 39 else 3 if n >= 7 else (n - 1) // 2
    p = (n - t) // 2
    q = p + t
    m = max(s + 1, n)
    T = lambda x: (x ^ (x >> 27)) & M
    for k in range(m):
        r1 = (1664525 * T(b[k % n] ^ b[(k + p) % n] ^ b[(k - 1) % n])) & M
        if k == 0:
            r2 = (r1 + s) & M
        elif k <= s:
            r2 = (r1 + k % n + v[k - 1]) & M
        else:
            r2 = (r1 + k % n) & M
        b[(k + p) % n] = (b[(k + p) % n] + r1) & M
        b[(k + q) % n] = (b[(k + q) % n] + r2) & M
        b[k % n] = r2
    for k in range(m, m + n):
        r3 = (1566083941 * T((b[k % n] + b[(k + p) % n] + b[(k - 1) % n]) & M)) & M
        r4 = (r3 - k % n) & M
        b[(k + p) % n] ^= r3
        b[(k + q) % n] ^= r4
        b[k % n] = r4
    return b


class Config:
    """scenes/*/tcnn_hairmsnn.json"""
    in_ch = 12
    out_ch = 3
    n_levels = 16
    feats = 2
    log2_hashmap = 15
    base_res = 16
    per_level_scale = 2.0
    blob_dims = 6
    blob_bins = 4
    width = 64
    padded_out = 16
    lr = 1e-2
    beta1 = 0.9
    beta2 = 0.99
    eps = 1e-15
    l2_reg = 1e-6
    decay_start = 4000
    decay_interval = 4000
    decay_base = 0.33
    loss_scale = 128.0
    seed = 1337

    def __init__(self, in_ch=12):
        self.in_ch = in_ch
        self.identity_dims = in_ch - 3 - self.blob_dims


def grid_layout(cfg):
    """GridEncodingTemplated ctor, encodings/grid.h:975-1015; grid_scale/resolution :195-204."""
    offsets, scales, ress = [0], [], []
    log2_pls = np.float32(np.log2(np.float32(cfg.per_level_scale)))
    for l in range(cfg.n_levels):
        scale = np.float32(np.exp2(np.float32(l) * log2_pls) * np.float32(cfg.base_res) - np.float32(1.0))
        res = int(np.ceil(scale)) + 1
        max_params = 0xFFFFFFFF // 2
        n = max_params if float(res) ** 3 > float(max_params) else res ** 3
        n = (n + 7) // 8 * 8
        n = min(n, 1 << cfg.log2_hashmap)
        offsets.append(offsets[-1] + n)
        scales.append(scale)
        ress.append(res)
    return offsets, scales, ress


def n_params(cfg):
    off, _, _ = grid_layout(cfg)
    n_matrix = cfg.width * cfg.width * 2 + cfg.padded_out * cfg.width
    return n_matrix + off[-1] * cfg.feats, n_matrix


def initial_params(cfg):
    """Trainer ctor + initialize_params (trainer.h:53-99): pcg32 seeded with the first word of
    std::seed_seq{1337}; FullyFusedMLP::initialize_params (src/fully_fused_mlp.cu:874-893) draws
    Xavier-uniform matrices in order (gpu_matrix.h:291-305); GridEncoding::initialize_params
    (encodings/grid.h:1357-1362) fills U(-1e-4, 1e-4) through generate_random_kernel
    (random.h:66-94: thread i draws elements i + n_threads*j from stream positions 4i + j)."""
    total, n_matrix = n_params(cfg)
    rng = Pcg32(seed_seq_generate([cfg.seed], 2)[0])
    out = np.zeros(total, np.float32)
    pos = 0
    for rows, cols in ((cfg.width, cfg.width), (cfg.width, cfg.width), (cfg.padded_out, cfg.width)):
        scale = np.float32(np.sqrt(np.float32(6.0) / np.float32(rows + cols)))
        f = rng.floats(rows * cols)
        out[pos:pos + rows * cols] = f * np.float32(2.0) * scale - scale
        pos += rows * cols
    n_grid = total - n_matrix
    n_thr = (n_grid + 3) // 4
    n_threads_total = (n_thr + 127) // 128 * 128
    f = rng.floats(4 * n_threads_total if 4 * n_threads_total < n_grid + 4 * 128 else n_grid + 4 * 128)
    i = np.arange(n_threads_total, dtype=np.int64)
    grid = np.zeros(n_grid, np.float32)
    for j in range(4):
        idx = i + n_threads_total * j
        ok = idx < n_grid
        src = 4 * i[ok] + j
        ok2 = src < len(f)
        # val * (upper - lower) + lower is contracted to one FMA by nvcc (random.h:66-94): a single rounding
        grid[idx[ok][ok2]] = (f[src[ok2]].astype(np.float64) * np.float64(np.float32(2e-4)) + np.float64(np.float32(-1e-4))).astype(np.float32)
    out[n_matrix:] = grid
    return out


def _h(x, half):
    return x.astype(np.float16).astype(np.float32) if half else x.astype(np.float32)


PRIMES = (np.uint32(1), np.uint32(2654435761), np.uint32(805459861))


def grid_index(pos_grid, res, hashmap_size):
    """grid_index<3, CoherentPrime> (encodings/grid.h:171-187, :127-131)."""
    pg = pos_grid.astype(np.uint32)
    stride = 1
    index = np.zeros(pg.shape[0], np.uint32)
    dim = 0
    while dim < 3 and stride <= hashmap_size:
        index = index + pg[:, dim] * np.uint32(stride & 0xFFFFFFFF)
        stride *= res
        dim += 1
    if hashmap_size < stride:
        index = (pg[:, 0] * PRIMES[0]) ^ (pg[:, 1] * PRIMES[1]) ^ (pg[:, 2] * PRIMES[2])
    return index % np.uint32(hashmap_size)


def grid_corners(cfg, x3, level):
    """pos_fract (common_device.h:434-445) + the 8 corner indices / weights (encodings/grid.h:330-345)."""
    off, scales, ress = grid_layout(cfg)
    scale, res = scales[level], ress[level]
    size = off[level + 1] - off[level]
    with np.errstate(over="ignore"):
        pos = (x3.astype(np.float32) * scale + np.float32(0.5)).astype(np.float32)
        fl = np.floor(pos)
        pg = fl.astype(np.int32).astype(np.uint32)
        fr = (pos - fl).astype(np.float32)
        idxs, ws = [], []
        for c in range(8):
            w = np.ones(x3.shape[0], np.float32)
            loc = pg.copy()
            for d in range(3):
                if (c >> d) & 1:
                    w = w * fr[:, d]
                    loc[:, d] = pg[:, d] + np.uint32(1)
                else:
                    w = w * (np.float32(1) - fr[:, d])
            idxs.append(grid_index(loc, res, size) + np.uint32(off[level]))
            ws.append(w.astype(np.float32))
    return idxs, ws


def quartic_cdf(x, inv_radius):
    """common_device.h:492-497"""
    u = (x * np.float32(inv_radius)).astype(np.float32)
    u2 = u * u
    u4 = u2 * u2
    v = np.float32(15.0 / 16.0) * u * (np.float32(1) - np.float32(2.0 / 3.0) * u2 + np.float32(1.0 / 5.0) * u4) + np.float32(0.5)
    return np.maximum(np.float32(0), np.minimum(np.float32(1), v)).astype(np.float32)


def encode(cfg, params, x, half=True, half_accumulate=True):
    """Composite[HashGrid | OneBlob | Identity] -> [N, 64] (encodings/composite.h:136-215,
    grid.h:221-351, oneblob.h:99-127, identity.h:46-66).  Unused trailing network inputs are
    filled with 1.0 by the last nested encoding."""
    x = np.asarray(x, np.float32)
    N = x.shape[0]
    _, n_matrix = n_params(cfg)
    table = _h(params[n_matrix:], half).reshape(-1, cfg.feats)
    out = np.ones((N, cfg.width), np.float32)
    for l in range(cfg.n_levels):
        idxs, ws = grid_corners(cfg, x[:, :3], l)
        if half and half_accumulate:
            acc = np.zeros((N, cfg.feats), np.float16)
            for i, w in zip(idxs, ws):
                acc = (acc + (w[:, None] * table[i]).astype(np.float16)).astype(np.float16)
            out[:, l * cfg.feats:(l + 1) * cfg.feats] = acc.astype(np.float32)
        else:
            acc = np.zeros((N, cfg.feats), np.float32)
            for i, w in zip(idxs, ws):
                acc += w[:, None] * table[i]
            out[:, l * cfg.feats:(l + 1) * cfg.feats] = _h(acc, half)
    c0 = cfg.n_levels * cfg.feats
    nb = cfg.blob_bins
    for j in range(cfg.blob_dims):
        xv = x[:, 3 + j]
        left = quartic_cdf(-xv, nb) + quartic_cdf(-xv - np.float32(1), nb) + quartic_cdf(-xv + np.float32(1), nb)
        for k in range(nb):
            rb = np.float32((k + 1) / nb)
            right = quartic_cdf(rb - xv, nb) + quartic_cdf(rb - xv - np.float32(1), nb) + quartic_cdf(rb - xv + np.float32(1), nb)
            out[:, c0 + j * nb + k] = _h(right - left, half)
            left = right
    c1 = c0 + cfg.blob_dims * nb
    for j in range(cfg.identity_dims):
        out[:, c1 + j] = _h(x[:, 3 + cfg.blob_dims + j], half)
    if cfg.identity_dims == 0:
        # render_nrc (9 inputs): the Identity encoding is dropped (composite.h:181), OneBlob is last and pads.
        # Its SoA padding writes the ones at offset N * n_dims_to_encode of its own slice (oneblob.h:221-225)
        # — rows 6..13 of the slice = network inputs 38..45, overwriting those OneBlob outputs — and never
        # touches the real padding rows 56..63, which stay whatever the allocation held (0 when fresh; pinned
        # by tests/golden/tcnn_9.npz).
        n_pad = cfg.width - c1
        out[:, c0 + cfg.blob_dims:c0 + cfg.blob_dims + n_pad] = 1.0
        out[:, c1:] = 0.0
    return out


def matrices(cfg, params, half=True):
    w = cfg.width
    W0 = _h(params[:w * w], half).reshape(w, w)
    W1 = _h(params[w * w:2 * w * w], half).reshape(w, w)
    Wo = _h(params[2 * w * w:2 * w * w + cfg.padded_out * w], half).reshape(cfg.padded_out, w)
    return W0, W1, Wo


def forward(cfg, params, x, half=True, keep=False):
    """kernel_mlp_fused (src/fully_fused_mlp.cu:499-557): h1 = ReLU(W0 e), h2 = ReLU(W1 h1),
    y = Wout h2; no biases; row-major [out][in] matrices; activations held in __half."""
    e = encode(cfg, params, x, half)
    W0, W1, Wo = matrices(cfg, params, half)
    h1 = _h(np.maximum(e @ W0.T, 0), half)
    h2 = _h(np.maximum(h1 @ W1.T, 0), half)
    y = _h(h2 @ Wo.T, half)
    if keep:
        return y, (e, h1, h2)
    return y[:, :cfg.out_ch].copy()


def loss_and_grad(cfg, y16, target, n_total_records=None, half=True):
    """relative_l2_luminance_loss (losses/relative_l2_luminance.h:40-87); stride 16, dims 3."""
    N = y16.shape[0]
    n_total = np.float32((n_total_records or N) * cfg.out_ch)
    pred = y16[:, :3].astype(np.float32)
    lum = np.float32(0.299) * pred[:, 0] + np.float32(0.587) * pred[:, 1] + np.float32(0.114) * pred[:, 2]
    denom = (lum * lum + np.float32(0.01)).astype(np.float32)
    diff = pred - np.asarray(target, np.float32)
    values = diff * diff / denom[:, None] / n_total
    grad = np.zeros_like(y16, dtype=np.float32)
    grad[:, :3] = _h(np.float32(cfg.loss_scale) * (np.float32(2) * diff / denom[:, None]) / n_total, half)
    return float(values.sum(dtype=np.float64)), grad


def backward(cfg, params, x, target, n_total_records=None, half=True):
    """Gradients (still multiplied by loss_scale) w.r.t. every parameter, tcnn order:
    kernel_mlp_fused_backward + weight-gradient GEMMs (src/fully_fused_mlp.cu:150-259,784-842),
    kernel_grid_backward (encodings/grid.h:395-516)."""
    y, (e, h1, h2) = forward(cfg, params, x, half, keep=True)
    loss, dy = loss_and_grad(cfg, y, target, n_total_records, half)
    W0, W1, Wo = matrices(cfg, params, half)
    dh2 = _h((dy @ Wo) * (h2 > 0), half)
    dh1 = _h((dh2 @ W1) * (h1 > 0), half)
    de = _h(dh1 @ W0, half)
    total, n_matrix = n_params(cfg)
    g = np.zeros(total, np.float32)
    w = cfg.width
    g[:w * w] = (dh1.T @ e).reshape(-1)
    g[w * w:2 * w * w] = (dh2.T @ h1).reshape(-1)
    g[2 * w * w:n_matrix] = (dy.T @ h2).reshape(-1)
    gt = np.zeros((total - n_matrix) // cfg.feats * cfg.feats, np.float32).reshape(-1, cfg.feats)
    xx = np.asarray(x, np.float32)
    for l in range(cfg.n_levels):
        idxs, ws = grid_corners(cfg, xx[:, :3], l)
        d = de[:, l * cfg.feats:(l + 1) * cfg.feats]
        for i, wgt in zip(idxs, ws):
            np.add.at(gt, i, _h(d * wgt[:, None], half))
    g[n_matrix:] = gt.reshape(-1)
    return loss, g


class Adam:
    """adam_step (optimizers/adam.h:48-120) under ExponentialDecayOptimizer::step
    (optimizers/exponential_decay.h:60-71)."""

    def __init__(self, cfg, params):
        self.cfg = cfg
        self.master = params.astype(np.float32).copy()
        self.m1 = np.zeros_like(self.master)
        self.m2 = np.zeros_like(self.master)
        self.steps = np.zeros(self.master.shape, np.uint32)
        self.step_count = 0
        self.lr_factor = np.float32(1.0)

    def step(self, grads_scaled, half=True):
        c = self.cfg
        _, n_matrix = n_params(c)
        if self.step_count == 0:
            self.lr_factor = np.float32(1.0)
        if self.step_count >= c.decay_start and (self.step_count - c.decay_start) % c.decay_interval == 0:
            self.lr_factor = np.float32(self.lr_factor * np.float32(c.decay_base))
        lr = np.float32(np.float32(c.lr) * self.lr_factor)
        self.step_count += 1
        g = (_h(grads_scaled, half) / np.float32(c.loss_scale)).astype(np.float32)
        is_matrix = np.arange(g.size) < n_matrix
        active = is_matrix | (g != 0)
        g = np.where(is_matrix, g + np.float32(c.l2_reg) * self.master, g).astype(np.float32)
        b1, b2 = np.float32(c.beta1), np.float32(c.beta2)
        m1 = (b1 * self.m1 + (np.float32(1) - b1) * g).astype(np.float32)
        m2 = (b2 * self.m2 + (np.float32(1) - b2) * g * g).astype(np.float32)
        steps = self.steps + np.uint32(1)
        sf = steps.astype(np.float32)
        lr_t = (lr * np.sqrt(np.float32(1) - np.power(b2, sf, dtype=np.float32)) / (np.float32(1) - np.power(b1, sf, dtype=np.float32))).astype(np.float32)
        eff = (lr_t / (np.sqrt(m2) + np.float32(c.eps))).astype(np.float32)
        new_w = (self.master - eff * m1).astype(np.float32)
        self.m1 = np.where(active, m1, self.m1)
        self.m2 = np.where(active, m2, self.m2)
        self.steps = np.where(active, steps, self.steps)
        self.master = np.where(active, new_w, self.master).astype(np.float32)
        return self.master
