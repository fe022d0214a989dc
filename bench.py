#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native HairMSNN per-path rendering loop.

Default workload (BASELINE.json metric / configs[3]): the reference's shipped scene `scenes/curly`
(50 000 strands, 3 391 580 Catmull-Rom segments, 78 520 head triangles, 4096x2048 RGBA32F environment +
1 directional light), loaded through hm_scene_load from assets/scenes/ (staged by scripts/stage_assets.py;
a procedural stand-in of the same size is used — and labelled — only when the staged files are missing),
render_hair_msnn at 1024x1024, BETA=1, MIS + ENV_PDF.  One "step" = one sample per pixel through the whole
frame loop: wavefront trace (G_BUFFER pass) -> online training step (16 384 records) -> MLP inference
(1 048 576 queries) -> composite (RENDER pass).  Metric: Mpaths/s = W*H*steps*n_gpus / seconds / 1e6.

--workload {msnn_b1 (default), msnn_b10, pt, nrc, straight4096} selects the other BASELINE configs; the
default single-GPU run also measures pt / nrc / msnn_b10 briefly (`other_workloads`) and the image gate
(relMSE of render_hair_msnn against a 500-spp render_path_tracing image, `image_gate`).

Multi-GPU (--gpus N under torchrun): one process per GPU.  The NCCL communicator lives in libhairmsnn.so
(hm_comm_*): attached to the renderer it shards samples (rank r renders sample ids r, r+N, ...: weak
scaling, per-GPU work fixed; straight4096: row bands), all-reduces the network's gradients inside every
training step and reduces framebuffers at the end of a job.  bench.py makes no collective of its own; the
128-byte communicator id travels through a file in /dev/shm.

`--impl reference` times the REFERENCE's own per-path code (cuda/hair_msnn.cu + headers compiled for the
host, oracle/_ref) plus the scalar network (oracle/mlp_scalar.c) on the box's CPU cores over a bounded,
frame-stratified sample of the same workload.
"""
import argparse
import os

# 10+ streams per renderer (main, order, one tail stream per frame in flight): more hardware queues than the
# default 8 avoid false serialisation between them (+1 % on B200); must be set before CUDA initialises
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import ctypes as C
import json
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

RECORDS = 16384
NODE_BYTES = 80          # 8-wide quantised BVH node (hm_bvh.h)
FLOPS_PER_QUERY = 16768           # 2*(64*64 + 64*64 + 64*3)
FLOPS_PER_RECORD = 50304          # forward + dgrad + wgrad
RELMSE_TOLERANCE = {"msnn_b1": 0.05, "msnn_b10": 0.01}   # gate against the 500-spp path-traced image

WORKLOADS = {
    # name: (scene, renderer kind, BETA, frame size or None = the scene file's)
    "msnn_b1": ("curly", "msnn", 1, None),
    "msnn_b10": ("curly", "msnn", 10, None),
    "pt": ("curly", "pt", 1, None),
    "nrc": ("curly", "nrc", 1, None),
    "straight4096": ("straight", "msnn", 1, 4096),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "bf16_tflops": d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0)),
                "bf16_tflops_burst": d.get("bf16_tflops", 1590.0), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "bf16_tflops_burst": 1590.0, "src": "fallback"}


class ClockSampler(threading.Thread):
    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.stop_flag = [], set(), False
        self.sm_max = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0])); self.sm_max = float(out[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), out[2:6]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["unsampled"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons)}


# ---- scenes -------------------------------------------------------------------------------------------------
def scene_config_path(scene, size=None):
    """Staged reference scene (assets/scenes/<scene>/config.json); a copy with another frame size is written
    next to it when `size` is given (paths inside resolve relative to the config's directory)."""
    base = os.path.join(ROOT, "assets", "scenes", scene, "config.json")
    if not os.path.exists(base) or not os.path.exists(os.path.join(ROOT, "assets", "scenes", "envmaps")):
        return None
    if size is None:
        return base
    cfg = json.load(open(base))
    cfg["integrator"]["width"] = cfg["integrator"]["height"] = int(size)
    out = os.path.join(os.path.dirname(base), f"config_{size}.json")
    tmp = out + f".{os.getpid()}.tmp"
    json.dump(cfg, open(tmp, "w"))
    os.replace(tmp, out)
    return out


def kw_from_config(path):
    """The reference-side parameters tests/refhost.py needs, read from a scene file (scene.cpp:119-339 keys)."""
    cfg = json.load(open(path))
    hair, integ, lights = cfg.get("hair", {}), cfg["integrator"], cfg.get("lights", {})
    dls = lights.get("directional", [])
    return dict(sigma_a=tuple(hair.get("sigma_a", (0.06, 0.1, 0.2))), beta_m=hair.get("beta_m", 0.3), beta_n=hair.get("beta_n", 0.3),
                alpha_deg=hair.get("alpha", 2.0), env_scale=lights.get("environment", {}).get("scale", 1.0), env_rotation=0.0,
                dl_from=[d["from"] for d in dls], dl_emit=[d["emit"] for d in dls], mis=integ.get("MIS", True),
                env_pdf=integ.get("ENV_PDF", True), path_v1=integ.get("path_v1", 1), path_v2=integ.get("path_v2", 40))


def make_scene(workload, strands=50000):
    """-> (api.Scene, reference-side kw, data label, W, H)"""
    from hairmsnn_b200 import api, synth
    scene, _, _, size = WORKLOADS[workload]
    cfgp = scene_config_path(scene, size)
    if cfgp:
        sc = api.Scene.load(cfgp)
        i = sc.info()
        return sc, kw_from_config(cfgp), f"reference-scene (scenes/{scene} of the reference, staged under assets/)", i.width, i.height
    W = H = size or 1024
    kw = synth.scene_kwargs(scene, W, H, num_strands=strands)
    return api.Scene.from_arrays(**kw), kw, f"synthetic (procedural stand-in for scenes/{scene}: staged assets missing)", W, H


def workload_text(workload, W, H, info):
    scene, kind, beta, _ = WORKLOADS[workload]
    what = {"msnn": f"render_hair_msnn BETA={beta} (online training 16384 records/step, {W * H} MLP queries/step)",
            "pt": "render_path_tracing (path_v2=40)", "nrc": "render_nrc (65536 training records/step, cache queries + training suffixes)"}[kind]
    return (f"{what} scenes/{scene} {W}x{H} ({info.num_strands} strands, {info.num_segments} segments, {info.num_triangles} triangles, "
            f"env {info.env_w}x{info.env_h} + {info.num_dlights} directional, MIS+ENV_PDF)")


# ---- CPU arm: the reference's per-path code + scalar network on the host cores --------------------------------
def scalar_mlp_lib():
    so = os.path.join(ROOT, "oracle", "_ref", "libmlp_scalar.so")
    src = os.path.join(ROOT, "oracle", "mlp_scalar.c")
    if not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-pthread", src, "-lm", "-o", so])
    lib = C.CDLL(so)
    lib.mlps_n_params.restype = C.c_size_t
    lib.mlps_gradients.restype = C.c_double
    return lib


class CpuArm:
    """One bounded sample of the msnn_b1 step on the host: 1/128 of everything a frame does.
    Rows y = 64 + 128 k (k = 0..H/128-1) — stratified over the whole frame, so the hair-hit fraction matches the
    GPU frame's — go through the reference's G_BUFFER pass (cuda/hair_msnn.cu compiled for the host); the rows'
    pixels are queried through the scalar network (inference of every pixel, as the reference does), their
    training records (W*rows/everyNth of the 16384) train it (forward + backward), and Adam updates 1/128 of the
    parameters.  Weights = the seeded initial weights, the same the GPU run starts from."""

    def __init__(self, sc, kw, W, H, beta_cli):
        from refhost import RefHost
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import mlp_oracle as mo
        self.ref = RefHost("msnn")
        self.ref.bind_all(sc, kw)
        self.W, self.H, self.beta = W, H, beta_cli - 1
        self.rows = list(range(64 % H, H, 128)) if H >= 128 else [H // 2]
        self.idxs = np.arange(RECORDS, dtype=np.int32)
        self.every_nth = W * H // RECORDS
        self.cores = os.cpu_count()
        self.lib = scalar_mlp_lib()
        self.params = mo.initial_params(mo.Config(12)).astype(np.float16).astype(np.float32)
        n = self.params.size
        self.master = self.params.copy()
        self.m1, self.m2, self.steps = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.uint32)
        self.grads = np.zeros(n, np.float32)
        self.paths_per_step = len(self.rows) * W
        self.slice = n // 128
        self.bufs = None

    def step(self, k):
        fp = C.POINTER(C.c_float)
        W = self.W
        self.bufs = self.ref.render_msnn_rows(k, W, self.H, self.beta, self.every_nth, self.idxs, self.rows, self.bufs, threads=self.cores)
        nn_in, tr_in, tr_out, _ = self.bufs
        # the sample's pixels and training records, gathered so that each network call spawns its threads once
        x = np.concatenate([nn_in[y * W:(y + 1) * W] for y in self.rows])
        out = np.empty((x.shape[0], 3), np.float32)
        self.lib.mlps_inference(self.params.ctypes.data_as(fp), x.ctypes.data_as(fp), x.shape[0], 12, out.ctypes.data_as(fp), self.cores)
        recs = np.concatenate([np.arange(y * W // self.every_nth, (y + 1) * W // self.every_nth) for y in self.rows])
        n_rec = len(recs)
        if n_rec:
            ti, to = np.ascontiguousarray(tr_in[recs]), np.ascontiguousarray(tr_out[recs])
            self.lib.mlps_gradients(self.params.ctypes.data_as(fp), ti.ctypes.data_as(fp), to.ctypes.data_as(fp), n_rec, 12, RECORDS,
                                    self.grads.ctypes.data_as(fp), min(self.cores, n_rec))
        first = (k % 128) * self.slice
        self.lib.mlps_adam(self.master.ctypes.data_as(fp), self.m1.ctypes.data_as(fp), self.m2.ctypes.data_as(fp),
                           self.steps.ctypes.data_as(C.POINTER(C.c_uint32)), self.grads.ctypes.data_as(fp), C.c_size_t(first), C.c_size_t(self.slice),
                           C.c_float(1e-2), C.c_float(0.9), C.c_float(0.99), C.c_float(1e-15), C.c_float(1e-6))
        return n_rec

    def sample_text(self, n_steps, secs):
        return (f"{n_steps} samples of rows {self.rows[0]}, {self.rows[0] + 128}, ... ({len(self.rows)} rows stratified over the frame, "
                f"{self.paths_per_step} paths each, {secs:.1f} s): G_BUFFER pass of the reference's cuda/hair_msnn.cu compiled for the host "
                f"(oracle/_ref; traversal = this repo's BVH + intersector on the host, OptiX being closed) + scalar network with the "
                f"seeded weights (oracle/mlp_scalar.c): inference of the rows' pixels, forward+backward of their {self.paths_per_step // self.every_nth} "
                f"training records, Adam on 1/128 of the parameters; {self.cores} threads")


def cpu_baseline(sc, kw, W, H, beta_cli, seconds_target=12.0):
    arm = CpuArm(sc, kw, W, H, beta_cli)
    t0 = time.perf_counter()
    n = 0
    while True:
        arm.step(n)
        n += 1
        if time.perf_counter() - t0 > seconds_target or n >= 4096:
            break
    dt = time.perf_counter() - t0
    return {"value": arm.paths_per_step * n / dt / 1e6, "unit": "Mpaths/s", "cores": arm.cores, "kind": "reference", "sample": arm.sample_text(n, dt)}


def run_reference(args, rank):
    if rank != 0:
        return
    if args.workload not in ("msnn_b1", "msnn_b10"):
        emit(json.dumps({"impl": "reference", "unavailable": f"the CPU arm implements the render_hair_msnn workloads, not {args.workload}"}))
        return
    sc, kw, data, W, H = make_scene(args.workload, args.strands)
    info = sc.info()
    arm = CpuArm(sc, kw, W, H, WORKLOADS[args.workload][2])
    times = []
    for step in range(args.warmup + args.steps):
        t = time.perf_counter()
        arm.step(step)
        dt = time.perf_counter() - t
        if step >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = arm.paths_per_step * len(times) / total / 1e6
    line = {"metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": data, "impl": "reference", "config": bench_config(args.workload, W, H, info, "spp"),
            "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": arm.cores, "kind": "reference", "sample": arm.sample_text(len(times), total)},
            "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(json.dumps(line))


def bench_config(workload, W, H, info, sharding):
    return {"workload": workload_text(workload, W, H, info), "name": workload,
            "l2": "working set (wide BVH nodes + leaf primitive copies ~3.5 GB, env tables 200 MB, path state ~200 MB per frame in flight) "
                  "exceeds the 126 MB L2; no flush needed",
            "sharding": sharding}


_REAL_STDOUT = None


def redirect_stdout():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(text):
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        print(text, flush=True)
    else:
        os.write(_REAL_STDOUT, (text + "\n").encode())


# ---- multi-GPU plumbing: communicator id through /dev/shm ---------------------------------------------------------
def exchange_comm_id(rank):
    from hairmsnn_b200 import api
    tag = f"{os.environ.get('MASTER_PORT', '0')}_{os.environ.get('TORCHELASTIC_RUN_ID', 'none')}_{os.getppid()}"
    path = os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp", f"hm_bench_comm_{tag}.id")
    if rank == 0:
        ident = api.Comm.unique_id()
        with open(path + ".tmp", "wb") as f:
            f.write(ident)
        os.replace(path + ".tmp", path)
        return ident, path
    for _ in range(3000):
        if os.path.exists(path):
            b = open(path, "rb").read()
            if len(b) == 128:
                return b, path
        time.sleep(0.1)
    raise RuntimeError("rank 0 never published the communicator id")


# ---- measurement of one renderer ------------------------------------------------------------------------------
def measure(r, api, torch, local_rank, comm, W, H, steps, warmup, peaks, kind, want_e2e=True, sampler=None):
    """Times `steps` frames of renderer `r` (device events on its stream, max over ranks), then the end-to-end
    leg, the per-stage table and the instrumented pass.  Returns a dict of raw results."""
    world = comm.world if comm else 1
    stream = torch.cuda.ExternalStream(r.stream, device=local_rank)

    def step():
        r.render_frames_async(1)

    def barrier():
        r.sync()
        torch.cuda.synchronize()
        if comm:
            comm.barrier()

    for _ in range(warmup):
        step()
    barrier()
    r.reset_stats()
    # event pairs around the launches the roofline line is about (main-piece k_trace) and the network's kernels;
    # timing all ~320 launches of a frame costs ~3 % of the frame rate, so the full per-stage table comes from a
    # second, untimed pass below
    r.set_profiling_stages((1 << 2) | (1 << 5) | (1 << 6))
    r.set_profiling_period(4)        # every 4th frame of the timed region carries the event pairs
    r.set_profiling(True)
    if sampler:
        sampler.start()
    launches0 = r.stats().kernel_launches
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        step()
    r.flush()            # the last frames' merged tail piece + order-stream work (held back until their group is full)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    if sampler:
        sampler.stop_flag = True
    st = r.stats()
    launches = st.kernel_launches - launches0
    r.set_profiling(False)
    r.set_profiling_period(1)
    if comm:
        ms = comm.all_reduce_max(ms)
    out = {"ms": ms, "value": W * H * steps * world / (ms * 1e-3) / 1e6, "launches": int(launches), "loss": st.last_loss}

    # second pass, all stages timed (not part of the headline number)
    r.reset_stats()
    r.set_profiling_stages(0xffffffff)
    r.set_profiling(True)
    for _ in range(steps):
        step()
    barrier()
    st_all = r.stats()
    r.set_profiling(False)

    if want_e2e:
        # End to end through the C ABI with host buffers, as a headless job runs: every step's 8-bit frame (the rows this
        # rank renders) is streamed to pinned host memory behind that step's composite — what the reference's viewer blits
        # per frame — and the job ends with the framebuffer reduction over the ranks (N > 1) and the read-back of the
        # fp32 average image (what saveEXR writes).  All of it is inside the timed region; two host buffer sets alternate.
        r0, r1 = r.rows()
        fb_host = [torch.empty((r1 - r0, W), dtype=torch.int32).pin_memory() for _ in range(2)]
        avg_host = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()
        barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            step()
            r.readback_rows_async(api.BUF_FB8, r0, r1 - r0, fb_host[i & 1].data_ptr())
        if comm:
            r.reduce_framebuffers()
        r.readback_async(api.BUF_FINAL_AVG, avg_host.data_ptr(), avg_host.numel() * 4)
        r.sync()
        e2e_s = time.perf_counter() - t0
        assert np.isfinite(avg_host.numpy()).all()
        if comm:
            e2e_s = comm.all_reduce_max(e2e_s)
        out["e2e"] = {"value": W * H * steps * world / e2e_s / 1e6, "unit": "Mpaths/s",
                      "h2d_bytes_per_step": int(api.lib.hm_frame_param_bytes() * launches / max(steps, 1)),
                      "d2h_bytes_per_step": int(fb_host[0].numel() * 4 + avg_host.numel() * 4 / max(steps, 1)),
                      "note": "per step: one frame through the C ABI (hm_render_frames_async; with N GPUs the gradient all-reduce is inside the call) "
                              "followed by hm_readback_rows_async of this rank's rows of the 8-bit frame into pinned host memory; once per job "
                              "(amortised over the steps in d2h_bytes_per_step): hm_reduce_framebuffers over the ranks and hm_readback_async of the "
                              "fp32 average image; host clock around the whole loop incl. the final sync; host->device traffic of a frame is its "
                              "kernel parameter blocks"}

    # instrumented pass: nodes visited / primitives tested per stage (SURVEY §8d per-ray bytes); every rank takes part
    r.reset_stats()
    r.set_collect_stats(True)
    n_inst = 2
    for _ in range(n_inst):
        step()
        r.sync()
    si = r.stats()
    r.set_collect_stats(False)
    barrier()

    # the timed region's event pairs sit on every 4th frame: scale its sums to all launches of the stage
    def scaled(ms_sum, stage):
        return ms_sum * st.stage_launches[stage] / max(st.timed_launches[stage], 1)
    stage_ms = {"primary": st_all.ms_primary, "shade": st_all.ms_shade, "trace": scaled(st.ms_extend, 2), "tail_piece": st_all.ms_shadow,
                "train": scaled(st.ms_train, 5), "infer": scaled(st.ms_infer, 6), "composite": st_all.ms_composite, "finalize": st_all.ms_finalize}
    stage_launches = dict(zip(("primary", "shade", "trace", "tail_piece", "finalize", "train", "infer", "composite"), st_all.stage_launches))
    stage_launches["trace"] = st.stage_launches[2]
    dominant = max(("primary", "trace"), key=lambda k: stage_ms[k])
    rays = {"primary": si.rays_primary, "trace": si.rays_extend + si.rays_shadow - si.rays_tail}
    nodes = {"primary": si.trav_nodes_primary, "trace": si.trav_nodes_extend + si.trav_nodes_shadow - si.trav_nodes_tail}
    prims = {"primary": si.trav_prims_primary, "trace": si.trav_prims_extend + si.trav_prims_shadow - si.trav_prims_tail}
    # algorithmic bytes per ray: 32 B ray + 16 B hit + 80 B per wide node visited + 64 B per primitive tested
    alg_bytes_per_step = (rays[dominant] * 48 + nodes[dominant] * NODE_BYTES + prims[dominant] * 64) / n_inst
    launches_dom = stage_launches[dominant] / steps
    avg_launch_ms = stage_ms[dominant] / max(stage_launches[dominant], 1)
    achieved = alg_bytes_per_step / max(launches_dom, 1) / max(avg_launch_ms * 1e-3, 1e-12) / 1e9
    traffic, traffic_note = None, None
    tp = os.path.join(ROOT, "profiles", f"k_{dominant}_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic, traffic_note = tj["traffic_bytes_per_launch"], tj["source"]
    out["roofline"] = {"kernel": f"k_{dominant}", "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                       "frac": achieved / peaks["hbm_gbs"], "peak_source": peaks["src"], "traffic": traffic, "traffic_note": traffic_note,
                       "algorithmic_bytes_per_step": alg_bytes_per_step,
                       "algorithmic_bytes_per_launch": alg_bytes_per_step / max(launches_dom, 1),
                       "rays_per_step": rays[dominant] / n_inst, "nodes_per_ray": nodes[dominant] / max(rays[dominant], 1),
                       "prims_per_ray": prims[dominant] / max(rays[dominant], 1), "avg_launch_ms": avg_launch_ms,
                       "launches_per_step": launches_dom,
                       "tail_piece": {"launches_per_step": stage_launches["tail_piece"] / steps, "rays_per_step": si.rays_tail / n_inst,
                                      "ms_per_step_sum": stage_ms["tail_piece"] / steps},
                       "timed_launches": int(st.timed_launches[2]) if dominant == "trace" else int(st_all.timed_launches[0]),
                       "note": "main-piece k_trace launches, CUDA events around the launches of every 4th frame inside the timed region (an event record "
                               "costs a few microseconds of launch gap: on every frame it costs 5 % of the frame rate); several frames are in flight, "
                               "so a launch shares the GPU with other frames' tail-piece and MLP kernels"}
    out["stage_ms_per_step"] = {k: v / steps for k, v in stage_ms.items()}
    out["rays_per_step"] = (si.rays_primary + si.rays_extend + si.rays_shadow) / n_inst
    if kind != "pt":
        in_ch, rows_submitted, records, _ = r.layout()
        rows_evaluated = rows_submitted
        if kind == "msnn":
            # rows the inference launch evaluates: 128-pixel tiles holding at least one hair hit (the RENDER pass reads no
            # other row's output; hm_renderer_set_skip_unused_queries).  Counted on the last frame's G-buffer.
            gflags = r.buffer(api.BUF_GBUFFER).reshape(-1, 4)[:, 3].copy().view(np.int32)
            hair_hit = ((gflags & 1) != 0) & ((gflags & 2) == 0)
            rows_evaluated = int(hair_hit.reshape(-1, 128).any(axis=1).sum()) * 128
        ms_infer = stage_ms["infer"] / steps
        ms_train = stage_ms["train"] / steps
        qps = rows_evaluated / (ms_infer * 1e-3) if ms_infer > 0 else None
        train_tflops = records * FLOPS_PER_RECORD / (ms_train * 1e-3) / 1e12 if ms_train > 0 else None
        out["mlp"] = {"queries_per_s": qps, "tflops": qps * FLOPS_PER_QUERY / 1e12 if qps else None,
                      "frac_of_tensor_peak": (qps * FLOPS_PER_QUERY / 1e12 / peaks["bf16_tflops"]) if qps else None,
                      "peak_tflops": peaks["bf16_tflops"], "ms_infer_per_step": ms_infer,
                      "rows_per_step_submitted": rows_submitted, "rows_per_step_evaluated": rows_evaluated,
                      "ms_train_per_step": ms_train, "train_records_per_step": records, "train_tflops": train_tflops,
                      "note": "in-frame numbers (the launches share the GPU with the next frames' traversal); the kernels timed alone: scripts/mlp_bench.py, profiles/"}
    return out


def mlp_alone(api, torch, local_rank, peaks):
    """The network's kernels timed alone (nothing else on the GPU), CUDA events on the network's stream: inference of 2^20
    queries on the synthetic inputs of SURVEY §8d (uniformly random positions: every gather is a distinct 32-byte L2 sector)
    and on pixel-coherent positions, and training steps of 16384 / 65536 records."""
    N = 1 << 20
    rng = np.random.default_rng(0)
    x = np.zeros((N, 12), np.float32)
    x[:, :3] = rng.uniform(-0.5, 0.5, (N, 3))
    d = rng.normal(size=(N, 6)).astype(np.float32)
    d[:, :3] /= np.linalg.norm(d[:, :3], axis=1, keepdims=True); d[:, 3:] /= np.linalg.norm(d[:, 3:], axis=1, keepdims=True)
    x[:, 3:9] = d
    xc = x.copy()
    g = np.arange(N)
    xc[:, 0] = ((g % 1024) / 1024.0 - 0.5) * 0.8; xc[:, 1] = ((g // 1024) / 1024.0 - 0.5) * 0.8
    xc[:, 2] = 0.1 * np.sin(xc[:, 0] * 9) * np.cos(xc[:, 1] * 7)
    dev = f"cuda:{local_rank}"
    m = api.Mlp.create(device=local_rank)
    xr, xco = torch.from_numpy(x).to(dev), torch.from_numpy(xc).to(dev)
    yo = torch.empty((N, 3), device=dev)
    st = torch.cuda.ExternalStream(m.stream, device=local_rank)
    torch.cuda.synchronize()

    def timed(fn, reps):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            fn()
        e1.record(st); e1.synchronize()
        return e0.elapsed_time(e1) / reps
    ms_r = timed(lambda: m.inference_device(xr.data_ptr(), yo.data_ptr(), N), 20)
    ms_c = timed(lambda: m.inference_device(xco.data_ptr(), yo.data_ptr(), N), 20)
    out = {"inference_ms_2p20_random_positions": ms_r, "inference_ms_2p20_pixel_coherent": ms_c,
           "queries_per_s_random": N / ms_r * 1e3, "queries_per_s_coherent": N / ms_c * 1e3,
           # 16 levels x 8 corners = 128 gathers per query, one 32-byte L2 sector each when positions are random
           "l2_gather_TBps_random": N * 128 * 32 / (ms_r * 1e-3) / 1e12,
           "l2_peak_TBps": 6300 * 1.965e9 / 1e12, "l2_peak_note": "LTS throughput cap ~6300 B/clk (B300_MICROARCH.md) x 1965 MHz",
           "tflops_random": N * FLOPS_PER_QUERY / (ms_r * 1e-3) / 1e12, "tensor_peak_tflops": peaks["bf16_tflops_burst"]}
    out["frac_of_l2_gather_peak_random"] = out["l2_gather_TBps_random"] / out["l2_peak_TBps"]
    out["frac_of_tensor_peak_random"] = out["tflops_random"] / peaks["bf16_tflops_burst"]
    for nrec in (16384, 65536):
        tx = xr[:nrec].contiguous(); ty = torch.rand((nrec, 3), device=dev)
        out[f"training_step_ms_{nrec}"] = timed(lambda: m.train_step_device(tx.data_ptr(), ty.data_ptr(), nrec), 30)
        fb = timed(lambda: m.forward_backward_device(tx.data_ptr(), ty.data_ptr(), nrec), 30)
        out[f"forward_backward_ms_{nrec}"] = fb
        out[f"forward_backward_tflops_{nrec}"] = nrec * FLOPS_PER_RECORD / (fb * 1e-3) / 1e12
    m.close()
    return out


def image_gate(sc, api, W, H, pt_spp, spp, pretrain):
    """relMSE of render_hair_msnn (BETA 1 and 10) against a render_path_tracing image of pt_spp samples."""
    def rel_mse(img, ref):
        img, ref = np.asarray(img, np.float64)[..., :3], np.asarray(ref, np.float64)[..., :3]
        return float(np.mean((img - ref) ** 2 / (ref ** 2 + 1e-2)))

    def render(kind, beta, n, offset=0):
        r = api.Renderer(sc, kind, beta_cli=beta, device=0)
        if offset:
            r.set_frame_schedule(offset, 1)
        if kind == api.HAIR_MSNN and pretrain:
            r.msnn_pretrain(pretrain)
        r.sync()
        t = time.perf_counter()
        done = 0
        while done < n:
            k = min(32, n - done)
            r.render_frames_async(k)
            done += k
        r.sync()
        dt = time.perf_counter() - t
        img = r.buffer(api.BUF_FINAL_AVG)
        r.close()
        return img, W * H * n / dt / 1e6

    gt, pt_rate = render(api.PATH_TRACING, 1, pt_spp)
    other, _ = render(api.PATH_TRACING, 1, spp, offset=100000)
    out = {"definition": "relMSE = mean over pixels and RGB of (I-R)^2/(R^2+0.01); R = render_path_tracing with pt_spp samples per pixel",
           "pt_spp": pt_spp, "spp": spp, "pretrain_steps": pretrain, "pt_mpaths_per_s": pt_rate,
           "relmse_pt_other_samples": rel_mse(other, gt), "tolerance": RELMSE_TOLERANCE}
    ok = True
    for name, beta in (("msnn_b1", 1), ("msnn_b10", 10)):
        img, rate = render(api.HAIR_MSNN, beta, spp)
        out[f"relmse_{name}"] = rel_mse(img, gt)
        out[f"mpaths_per_s_{name}"] = rate
        ok = ok and out[f"relmse_{name}"] <= RELMSE_TOLERANCE[name]
    out["pass"] = bool(ok)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="msnn_b1", choices=sorted(WORKLOADS))
    ap.add_argument("--strands", type=int, default=50000, help="synthetic fallback only")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the brief pt / nrc / msnn_b10 measurements of the default run")
    ap.add_argument("--no-gate", action="store_true", help="skip the image gate of the default run")
    ap.add_argument("--gate-spp", type=int, default=500)
    ap.add_argument("--beta-sweep", default="", help="comma-separated BETA values measured after the main run (render_hair_msnn workloads; BASELINE config 5)")
    ap.add_argument("--size", type=int, default=0, help="override the frame size of the workload (testing the band path on fewer GPUs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout carries exactly one JSON line (rank 0): libraries that write to fd 1 (NCCL prints its
    # version banner there) are sent to stderr for the duration of the run
    redirect_stdout()
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    from hairmsnn_b200 import api

    torch.cuda.set_device(local_rank)
    peaks = load_peaks()
    if args.size:
        w = WORKLOADS[args.workload]
        WORKLOADS[args.workload] = (w[0], w[1], w[2], args.size)
    scene_name, kind, beta_cli, size = WORKLOADS[args.workload]
    bands = args.workload == "straight4096"
    if bands and (size * size) // world > 2048 * 2048:
        raise SystemExit("straight4096 renders 4096x4096 on row bands of at most 2048x2048 pixels: needs --gpus 4 or 8")

    comm, id_path = None, None
    if world > 1:
        ident, id_path = exchange_comm_id(rank)
        comm = api.Comm(ident, rank, world, local_rank)

    t0 = time.time()
    if world == 1:
        sc, kw, data, W, H = make_scene(args.workload, args.strands)
    else:
        # one rank builds the acceleration structure (and keeps the binary tree the CPU arm needs), the others
        # restore the GPU-side tree from a RAM-backed cache instead of N concurrent builds
        cache = os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp", f"hm_bvh_cache_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}")
        os.makedirs(cache, exist_ok=True)
        if rank == 0:
            sc, kw, data, W, H = make_scene(args.workload, args.strands)
            sc.save_bvh_cache(cache)
        comm.barrier()
        if rank != 0:
            os.environ["HM_BVH_CACHE"] = cache
            sc, kw, data, W, H = make_scene(args.workload, args.strands)
            del os.environ["HM_BVH_CACHE"]
        comm.barrier()
        if rank == 0:
            import shutil
            shutil.rmtree(cache, ignore_errors=True)
            try:
                os.remove(id_path)
            except OSError:
                pass
    info = sc.info()
    KIND = {"msnn": api.HAIR_MSNN, "pt": api.PATH_TRACING, "nrc": api.NRC}
    r = api.Renderer(sc, KIND[kind], beta_cli=beta_cli, device=local_rank, rank=rank if bands else 0, world=world if bands else 1)
    if comm:
        r.set_comm(comm)      # sample groups (or row bands) + gradient all-reduce inside every training step
    log(f"[rank {rank}] scene + renderer ready in {time.time() - t0:.1f}s ({data})")

    sampler = ClockSampler(local_rank)
    # row bands split ONE frame over the ranks: a step is one frame of the job (strong scaling in N)
    res = measure(r, api, torch, local_rank, comm, W, H, args.steps, args.warmup, peaks, kind, sampler=sampler)
    if bands:
        res["value"] /= world
        res["e2e"]["value"] /= world
    if comm:
        r.reduce_framebuffers()     # the job's one framebuffer reduction (not part of a step)
    sweep = None
    if args.beta_sweep and kind == "msnn":
        # BASELINE config 5: the same frame at other BETA values (every rank takes part: the steps hold collectives)
        sweep = {}
        r.close()
        for b in [int(x) for x in args.beta_sweep.split(",") if x]:
            rb = api.Renderer(sc, KIND[kind], beta_cli=b, device=local_rank, rank=rank if bands else 0, world=world if bands else 1)
            if comm:
                rb.set_comm(comm)
            m = measure(rb, api, torch, local_rank, comm, W, H, args.steps, args.warmup, peaks, kind, want_e2e=False)
            rb.close()
            v = m["value"] / world if bands else m["value"]
            sweep[str(b)] = {"value": v, "unit": "Mpaths/s", "ms_per_step": m["ms"] / args.steps, "rays_per_step": m["rays_per_step"]}
            if rank == 0:
                log(f"[beta sweep] BETA={b}: {v:.1f} Mpaths/s")
    if rank != 0:
        r.close()
        comm.barrier()
        comm.close()
        return

    others, gate, cb = None, None, None
    default_single = args.workload == "msnn_b1" and world == 1
    r.close()
    if default_single and not args.no_others:
        others = {}
        for name in ("pt", "nrc", "msnn_b10"):
            _, k2, b2, _ = WORKLOADS[name]
            r2 = api.Renderer(sc, KIND[k2], beta_cli=b2, device=local_rank)
            m = measure(r2, api, torch, local_rank, None, W, H, 8, 3, peaks, k2, want_e2e=False)
            r2.close()
            others[name] = {"workload": workload_text(name, W, H, info), "value": m["value"], "unit": "Mpaths/s", "ms_per_step": m["ms"] / 8, "steps": 8,
                            "warmup": 3, "gpu_launches": m["launches"], "rays_per_step": m["rays_per_step"], "roofline": m["roofline"],
                            "mlp": m.get("mlp"), "stage_ms_per_step": m["stage_ms_per_step"]}
            log(f"[other workload] {name}: {m['value']:.1f} Mpaths/s")
    if default_single and not args.no_gate and data.startswith("reference-scene"):
        gate = image_gate(sc, api, W, H, args.gate_spp, args.gate_spp, 200)
        log(f"[image gate] {gate}")
    if not args.no_cpu_baseline and kind == "msnn":
        cb = cpu_baseline(sc, kw, W, H, beta_cli)
    if default_single and res.get("mlp") is not None:
        res["mlp"]["alone"] = mlp_alone(api, torch, local_rank, peaks)

    line = {"metric": "Mpaths/s", "value": res["value"], "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": res["ms"] / args.steps, "higher_is_better": True, "scaling": "strong" if bands else "weak", "vs_baseline": None,
            "dtype": "f32", "data": data, "config": bench_config(args.workload, W, H, info, "row bands" if bands else "spp"),
            "clocks": sampler.summary(), "e2e": res["e2e"], "gpu_launches": res["launches"],
            "roofline": res["roofline"], "mlp": res.get("mlp"), "cpu_baseline": cb,
            "stage_ms_per_step": res["stage_ms_per_step"], "training_loss": res["loss"],
            "collectives": ("ncclAllReduce of 1000448 fp32 gradients per training step inside hm_render_frames (libhairmsnn.so); "
                            "bench.py issues none") if world > 1 else None,
            "image_gate": gate, "other_workloads": others, "beta_sweep": sweep}
    emit(json.dumps(line))
    if comm:
        comm.barrier()
        comm.close()


if __name__ == "__main__":
    main()
