#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native HairMSNN per-path rendering loop.

Workload (BASELINE.json metric / configs[3]): synthetic stand-in for scenes/curly
(50 000 strands, 3.4 M Catmull-Rom segments, ~78 k head triangles, 4096x2048 RGBA32F
environment + 1 directional light), render_hair_msnn at 1024x1024, BETA=1, MIS + ENV_PDF.
One "step" = one sample per pixel through the whole frame loop: wavefront trace (G_BUFFER
pass) -> online training step (16 384 records) -> MLP inference (1 048 576 queries) ->
composite (RENDER pass).  Metric: Mpaths/s = W*H*steps*n_gpus / seconds / 1e6.

Multi-GPU (--gpus N under torchrun): samples are sharded (rank r renders sample indices
r, r+N, ...: weak scaling, per-GPU work fixed), MLP gradients are all-reduced over NCCL
every step so all replicas hold identical weights; framebuffers would be summed once at the
end of a job (not part of a step).

`--impl reference` times the REFERENCE's own per-path code (cuda/hair_msnn.cu + headers,
compiled for the host in oracle/_ref) on the box's CPU cores over a bounded band of the
same frame.
"""
import argparse
import os

# 10+ streams per renderer (main, order, one tail stream per frame in flight): more hardware queues than the
# default 8 avoid false serialisation between them (+1 % on B200); must be set before CUDA initialises
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

W = H = 1024
BETA_CLI = 1
RECORDS = 16384
NODE_BYTES = 80          # 8-wide quantised BVH node (hm_bvh.h)
FLOPS_PER_QUERY = 16768           # 2*(64*64 + 64*64 + 64*3)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "bf16_tflops": d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0)), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "src": "fallback"}


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


def make_scene(num_strands):
    from hairmsnn_b200 import api, synth
    kw = synth.scene_kwargs("curly", W, H, num_strands=num_strands)
    sc = api.Scene.from_arrays(**kw)
    return sc, kw


CONFIG = {"workload": "render_hair_msnn synthetic-curly 1024x1024 BETA=1 (50k strands, 3.4M segments, env 4096x2048 + 1 directional, MIS+ENV_PDF, online training 16384 records/step, 1048576 MLP queries/step)",
          "l2": "working set (wide BVH nodes 1.1 GB + leaf primitive copies 2.8 GB + control points 57 MB + env tables 200 MB + path state 180 MB per frame in flight) exceeds the 126 MB L2; no flush needed",
          "sharding": "spp"}


def reference_band():
    """Rows of the frame the CPU arm renders per step: a band through the hair volume."""
    return H // 2 - 4, H // 2 + 4


def run_reference(args, rank):
    if rank != 0:
        return
    from refhost import RefHost
    sc, kw = make_scene(args.strands)
    ref = RefHost("msnn")
    ref.bind_all(sc, kw)
    y0, y1 = reference_band()
    idxs = np.arange(RECORDS, dtype=np.int32)
    cores = os.cpu_count()
    every_nth = W * H // RECORDS
    times = []
    for step in range(args.warmup + args.steps):
        t = time.perf_counter()
        ref.render_msnn_gbuffer(step, W, H, BETA_CLI - 1, every_nth, idxs, y0=y0, y1=y1, threads=cores)
        dt = time.perf_counter() - t
        if step >= args.warmup:
            times.append(dt)
    paths = (y1 - y0) * W
    total = sum(times)
    value = paths * len(times) / total / 1e6
    line = {"metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference", "config": CONFIG,
            "cpu_baseline": {"value": value, "unit": "Mpaths/s", "cores": cores, "kind": "reference",
                             "sample": f"rows {y0}..{y1 - 1} of the 1024x1024 frame ({paths} paths) per step: G_BUFFER pass of cuda/hair_msnn.cu compiled for the host, all host threads"},
            "e2e": {"value": value, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(json.dumps(line))


def cpu_baseline(sc, kw, seconds_target=12.0):
    from refhost import RefHost
    ref = RefHost("msnn")
    ref.bind_all(sc, kw)
    y0, y1 = reference_band()
    idxs = np.arange(RECORDS, dtype=np.int32)
    cores = os.cpu_count()
    every_nth = W * H // RECORDS
    t0 = time.perf_counter()
    n = 0
    while True:
        ref.render_msnn_gbuffer(n, W, H, BETA_CLI - 1, every_nth, idxs, y0=y0, y1=y1, threads=cores)
        n += 1
        if time.perf_counter() - t0 > seconds_target or n >= 4096:
            break
    dt = time.perf_counter() - t0
    paths = (y1 - y0) * W * n
    return {"value": paths / dt / 1e6, "unit": "Mpaths/s", "cores": cores, "kind": "reference",
            "sample": f"{n} samples of rows {y0}..{y1 - 1} ({paths} paths, {dt:.1f} s): the reference's hair_msnn.cu G_BUFFER pass compiled for the host (oracle/_ref), {cores} threads"}


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--strands", type=int, default=50000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    import torch.distributed as dist
    from hairmsnn_b200 import api

    torch.cuda.set_device(local_rank)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=180))
    peaks = load_peaks()

    t0 = time.time()
    if world == 1:
        sc, kw = make_scene(args.strands)
    else:
        # one rank builds the acceleration structure (and keeps the binary tree the CPU arm needs), the others
        # restore the GPU-side tree from a RAM-backed cache instead of N concurrent 40-second builds
        cache = os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp", f"hm_bvh_cache_{os.environ.get('MASTER_PORT', '0')}")
        os.makedirs(cache, exist_ok=True)
        if rank == 0:
            sc, kw = make_scene(args.strands)
            sc.save_bvh_cache(cache)
        dist.barrier()
        if rank != 0:
            os.environ["HM_BVH_CACHE"] = cache
            sc, kw = make_scene(args.strands)
            del os.environ["HM_BVH_CACHE"]
        dist.barrier()
        if rank == 0:
            import shutil
            shutil.rmtree(cache, ignore_errors=True)
    r = api.Renderer(sc, api.HAIR_MSNN, beta_cli=BETA_CLI, device=local_rank)
    r.set_frame_schedule(rank, world)
    mlp = r.mlp()
    log(f"[rank {rank}] scene + renderer ready in {time.time() - t0:.1f}s")
    stream = torch.cuda.ExternalStream(r.stream, device=local_rank)

    class DevBuf:
        def __init__(self, ptr, nbytes, dtype, shape):
            self.__cuda_array_interface__ = {"shape": shape, "typestr": dtype, "data": (ptr, False), "version": 3}
    gptr, gcount = mlp.gradients_device()
    grads = torch.as_tensor(DevBuf(gptr, gcount * 4, "<f4", (gcount,)), device=f"cuda:{local_rank}")
    tin_ptr, _ = r.device_buffer(api.BUF_NN_TRAIN_INPUT)
    tout_ptr, _ = r.device_buffer(api.BUF_NN_TRAIN_OUTPUT)

    def step():
        if world == 1:
            r.render_frames_async(1)
            return
        # split frame: trace -> backward -> gradient all-reduce over NVLink -> Adam -> inference + composite
        r.msnn_trace()
        mlp.forward_backward_device(tin_ptr, tout_ptr, RECORDS, RECORDS * world)
        with torch.cuda.stream(stream):
            dist.all_reduce(grads)
        r.msnn_train_apply()
        r.msnn_finish()

    def barrier():
        r.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        r.sync()

    for _ in range(args.warmup):
        step()
    barrier()

    # ---- timed region: device-resident -------------------------------------------------
    r.reset_stats()
    # event pairs around the launches the roofline line is about (main-piece k_trace, 2 per frame) and the
    # network inference; timing all ~320 launches of a frame costs ~3 % of the frame rate, so the full
    # per-stage table comes from a second, untimed pass below
    r.set_profiling_stages((1 << 2) | (1 << 6))
    r.set_profiling(True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = r.stats().kernel_launches
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    sampler.stop_flag = True
    st = r.stats()
    launches = st.kernel_launches - launches0
    r.set_profiling(False)
    # second pass, all stages timed (not part of the headline number)
    r.reset_stats()
    r.set_profiling_stages(0xffffffff)
    r.set_profiling(True)
    for _ in range(args.steps):
        step()
    barrier()
    st_all = r.stats()
    r.set_profiling(False)
    if world > 1:
        t = torch.tensor([ms], device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = W * H * args.steps * world / (ms * 1e-3) / 1e6

    # ---- end to end through the C ABI with host buffers ---------------------------------
    # every step's results (8-bit framebuffer + fp32 average) are streamed to pinned host memory
    # behind that step's composite; two host buffer sets alternate
    fb_host = [torch.empty((H, W), dtype=torch.int32).pin_memory() for _ in range(2)]
    avg_host = [torch.empty((H, W, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step()
        r.readback_async(api.BUF_FB8, fb_host[i & 1].data_ptr(), fb_host[0].numel() * 4)
        r.readback_async(api.BUF_FINAL_AVG, avg_host[i & 1].data_ptr(), avg_host[0].numel() * 4)
    r.sync()
    e2e_s = time.perf_counter() - t0
    assert np.isfinite(avg_host[(args.steps - 1) & 1].numpy()).all()
    if world > 1:
        t = torch.tensor([e2e_s], device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = W * H * args.steps * world / e2e_s / 1e6
    frame_param_bytes = int(api.lib.hm_frame_param_bytes())
    launches_per_step = launches / max(args.steps, 1)

    # ---- roofline of the dominant kernel -------------------------------------------------
    # instrumented pass: nodes visited / primitives tested per stage (SURVEY §8d per-ray bytes).
    # Every rank takes part (the step holds a collective when world > 1); rank 0 reports.
    r.reset_stats()
    r.set_collect_stats(True)
    n_inst = 2
    for _ in range(n_inst):
        step()
        r.sync()
    si = r.stats()
    r.set_collect_stats(False)
    if rank != 0:
        barrier()
        dist.destroy_process_group()
        return
    if world > 1:
        barrier()
    # k_trace: one launch per path vertex tracing its occlusion probes and continuation rays.  A frame's
    # launches split into the main piece (vertices 0..BETA of every path: ~97% of the secondary rays, on
    # the main stream) and the tail piece (the few training paths' deeper vertices: dozens of tiny
    # latency-bound launches on a side stream).  The roofline line is about the main-piece launches.
    stage_ms = {"primary": st_all.ms_primary, "shade": st_all.ms_shade, "trace": st.ms_extend, "tail_piece": st_all.ms_shadow,
                "train": st_all.ms_train, "infer": st.ms_infer, "composite": st_all.ms_composite, "finalize": st_all.ms_finalize}
    stage_launches = dict(zip(("primary", "shade", "trace", "tail_piece", "finalize", "train", "infer", "composite"), st_all.stage_launches))
    stage_launches["trace"] = st.stage_launches[2]
    dominant = max(("primary", "trace"), key=lambda k: stage_ms[k])
    rays = {"primary": si.rays_primary, "trace": si.rays_extend + si.rays_shadow - si.rays_tail}
    nodes = {"primary": si.trav_nodes_primary, "trace": si.trav_nodes_extend + si.trav_nodes_shadow - si.trav_nodes_tail}
    prims = {"primary": si.trav_prims_primary, "trace": si.trav_prims_extend + si.trav_prims_shadow - si.trav_prims_tail}
    # algorithmic bytes per ray: 32 B ray + 16 B hit + 80 B per wide node visited + 64 B per primitive tested
    alg_bytes_per_step = (rays[dominant] * 48 + nodes[dominant] * NODE_BYTES + prims[dominant] * 64) / n_inst
    launches_dom = stage_launches[dominant] / args.steps
    avg_launch_ms = stage_ms[dominant] / max(stage_launches[dominant], 1)
    achieved = alg_bytes_per_step / max(launches_dom, 1) / (avg_launch_ms * 1e-3) / 1e9
    # DRAM traffic of the dominant kernel from the committed ncu --set full capture (profiles/)
    traffic, traffic_note = None, None
    tp = os.path.join(ROOT, "profiles", f"k_{dominant}_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic = tj["traffic_bytes_per_launch"]
        traffic_note = tj["source"]
    roofline = {"kernel": f"k_{dominant}", "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "peak_source": peaks["src"], "traffic": traffic, "traffic_note": traffic_note,
                "algorithmic_bytes_per_step": alg_bytes_per_step,
                "algorithmic_bytes_per_launch": alg_bytes_per_step / max(launches_dom, 1),
                "rays_per_step": rays[dominant] / n_inst, "nodes_per_ray": nodes[dominant] / max(rays[dominant], 1),
                "prims_per_ray": prims[dominant] / max(rays[dominant], 1), "avg_launch_ms": avg_launch_ms,
                "launches_per_step": launches_dom,
                "tail_piece": {"launches_per_step": stage_launches["tail_piece"] / args.steps, "rays_per_step": si.rays_tail / n_inst,
                               "ms_per_step_sum": stage_ms["tail_piece"] / args.steps},
                "note": "main-piece k_trace launches (2 per frame at BETA=1), CUDA events around each launch inside the timed region; "
                        "8 frames are in flight, so a launch shares the GPU with other frames' tail-piece and MLP kernels; the kernel is "
                        "bound by dependent-fetch latency and SIMT divergence, not by bandwidth (profiles/)"}
    # rows the inference launch evaluates: 128-pixel tiles holding at least one hair hit (the RENDER pass reads no
    # other row's output; hm_renderer_set_skip_unused_queries).  Counted on the last frame's G-buffer.
    gflags = r.buffer(api.BUF_GBUFFER).reshape(-1, 4)[:, 3].copy().view(np.int32)
    hair_hit = ((gflags & 1) != 0) & ((gflags & 2) == 0)
    rows_evaluated = int(hair_hit.reshape(-1, 128).any(axis=1).sum()) * 128
    mlp_qps = rows_evaluated / (stage_ms["infer"] / args.steps * 1e-3) if stage_ms["infer"] > 0 else None
    mlp_info = {"queries_per_s": mlp_qps, "tflops": mlp_qps * FLOPS_PER_QUERY / 1e12 if mlp_qps else None,
                "frac_of_tensor_peak": (mlp_qps * FLOPS_PER_QUERY / 1e12 / peaks["bf16_tflops"]) if mlp_qps else None,
                "peak_tflops": peaks["bf16_tflops"], "ms_infer_per_step": stage_ms["infer"] / args.steps,
                "rows_per_step_submitted": W * H, "rows_per_step_evaluated": rows_evaluated,
                "ms_train_per_step": stage_ms["train"] / args.steps}

    cb = None if args.no_cpu_baseline else cpu_baseline(sc, kw)

    line = {"metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": CONFIG,
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": int(frame_param_bytes * launches_per_step),
                    "d2h_bytes_per_step": int(fb_host[0].numel() * 4 + avg_host[0].numel() * 4),
                    "note": "per step: one frame through the C ABI (hm_render_frames_async / split-frame calls) followed by hm_readback_async of the 8-bit framebuffer and the fp32 average buffer into pinned host memory, host clock around the whole loop incl. the final sync; host->device traffic of a frame is its kernel parameter blocks"},
            "gpu_launches": int(launches),
            "roofline": roofline, "mlp": mlp_info, "cpu_baseline": cb,
            "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()},
            "training_loss": st.last_loss}
    emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
