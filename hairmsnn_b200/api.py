"""ctypes binding of the C ABI in include/hairmsnn.h.

Host-side mirror of the reference's entry points for Python callers (tests, bench.py):
`Scene` ~ parseScene + Scene (scene.cpp), `Renderer` ~ RenderWindowPT /
RenderWindow_HairMSNN (render_*.cu), `Mlp` ~ TINY_MLP (cuda/neural_network.cu).
There is no fallback: if the CUDA library is missing, importing this module raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HM_LIB") or os.path.join(_HERE, "lib", "libhairmsnn.so")   # HM_LIB: A/B builds of the same library

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `make` (or __graft_entry__.build()). "
        "hairmsnn_b200 has no CPU or PyTorch fallback."
    )
lib = C.CDLL(LIB_PATH)

PATH_TRACING, NRC, HAIR_MSNN = 0, 1, 2
BUF_FINAL_AVG, BUF_FINAL_ACCUM, BUF_PT_AVG, BUF_PT_ACCUM, BUF_NN_AVG, BUF_NN_ACCUM, BUF_FB8 = range(7)
BUF_NN_FRAME_INPUT, BUF_NN_FRAME_OUTPUT, BUF_NN_TRAIN_INPUT, BUF_NN_TRAIN_OUTPUT, BUF_GBUFFER, BUF_TRAIN_IDXS = range(7, 13)
BUF_GBUFFER_B, BUF_NRC_TRAIN_RECORDS, BUF_SCENE_INDICES, BUF_SCENE_POINTS = 13, 14, 15, 16
BUF_ENV_CPDF, BUF_ENV_CCDF, BUF_ENV_MPDF, BUF_ENV_MCDF = 17, 18, 19, 20
NRC_MAX_BOUNCES = 40

_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)


class SceneDesc(C.Structure):
    _fields_ = [
        ("control_points", _fp), ("num_control_points", C.c_int),
        ("segment_first_cp", _ip), ("num_segments", C.c_int), ("num_strands", C.c_int),
        ("hair_min", C.c_float * 3), ("hair_max", C.c_float * 3),
        ("tri_vertices", _fp), ("tri_normals", _fp), ("num_triangles", C.c_int),
        ("surface_kd", C.c_float * 3), ("surface_alpha", C.c_float),
        ("cam_from", C.c_float * 3), ("cam_to", C.c_float * 3), ("cam_up", C.c_float * 3), ("cos_fovy", C.c_float),
        ("sigma_a", C.c_float * 3), ("beta_m", C.c_float), ("beta_n", C.c_float), ("alpha", C.c_float),
        ("gains", C.c_float * 4),
        ("env_rgba", _fp), ("env_w", C.c_int), ("env_h", C.c_int), ("env_scale", C.c_float), ("env_rotation", C.c_float),
        ("dl_from", _fp), ("dl_emit", _fp), ("num_dlights", C.c_int),
        ("width", C.c_int), ("height", C.c_int), ("spp", C.c_int), ("path_v1", C.c_int), ("path_v2", C.c_int),
        ("mis", C.c_int), ("env_pdf", C.c_int),
        ("tcnn_config_path", C.c_char_p),
    ]


class SceneInfo(C.Structure):
    _fields_ = [
        ("width", C.c_int), ("height", C.c_int), ("spp", C.c_int), ("path_v1", C.c_int), ("path_v2", C.c_int),
        ("num_segments", C.c_int), ("num_control_points", C.c_int), ("num_triangles", C.c_int),
        ("num_strands", C.c_int), ("num_bvh_nodes", C.c_int),
        ("scene_scale", C.c_float),
        ("cam_pos", C.c_float * 3), ("cam_d00", C.c_float * 3), ("cam_du", C.c_float * 3), ("cam_dv", C.c_float * 3),
        ("env_w", C.c_int), ("env_h", C.c_int), ("num_dlights", C.c_int),
        ("num_wide_nodes", C.c_int), ("num_wide_leaf_refs", C.c_int), ("wide_depth", C.c_int),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("ms_primary", C.c_double), ("ms_shade", C.c_double), ("ms_extend", C.c_double), ("ms_shadow", C.c_double),
        ("ms_finalize", C.c_double), ("ms_train", C.c_double), ("ms_infer", C.c_double), ("ms_composite", C.c_double),
        ("ms_total", C.c_double),
        ("rays_primary", C.c_uint64), ("rays_extend", C.c_uint64), ("rays_shadow", C.c_uint64), ("shade_items", C.c_uint64),
        ("kernel_launches", C.c_uint64),
        ("stage_launches", C.c_uint64 * 8),
        ("trav_nodes_extend", C.c_uint64), ("trav_prims_extend", C.c_uint64), ("trav_nodes_shadow", C.c_uint64),
        ("trav_prims_shadow", C.c_uint64), ("trav_nodes_primary", C.c_uint64), ("trav_prims_primary", C.c_uint64),
        ("trav_nodes_tail", C.c_uint64), ("trav_prims_tail", C.c_uint64), ("rays_tail", C.c_uint64),
        ("last_loss", C.c_float), ("frames", C.c_int),
        ("timed_launches", C.c_uint64 * 8),
    ]


lib.hm_last_error.restype = C.c_char_p
lib.hm_renderer_stream.restype = C.c_void_p
lib.hm_renderer_mlp.restype = C.c_void_p
lib.hm_mlp_stream.restype = C.c_void_p
lib.hm_mlp_n_params.restype = C.c_size_t
lib.hm_mlp_launch_count.restype = C.c_uint64
for _name in ("hm_renderer_stream", "hm_renderer_mlp", "hm_renderer_destroy", "hm_render_frames", "hm_render_frames_async", "hm_render_flush",
              "hm_renderer_sync", "hm_renderer_reset_accumulation", "hm_renderer_accum_id", "hm_msnn_trace",
              "hm_msnn_train_backward", "hm_msnn_train_apply", "hm_msnn_finish", "hm_nrc_trace", "hm_nrc_query",
              "hm_nrc_train_backward", "hm_nrc_train_apply", "hm_nrc_end", "hm_mlp_stream", "hm_mlp_n_params",
              "hm_mlp_launch_count", "hm_mlp_destroy", "hm_mlp_optimizer_step", "hm_mlp_reset", "hm_mlp_reinitialize",
              "hm_scene_free", "hm_comm_destroy", "hm_comm_barrier", "hm_reduce_framebuffers"):
    getattr(lib, _name).argtypes = [C.c_void_p]


lib.hm_renderer_set_comm.argtypes = [C.c_void_p, C.c_void_p]
lib.hm_comm_all_reduce_max.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
lib.hm_comm_all_reduce_sum.argtypes = [C.c_void_p, C.POINTER(C.c_double)]


class HairMSNNError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code


def _check(code):
    if code != 0:
        raise HairMSNNError(code, lib.hm_last_error().decode("utf-8", "replace"))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a, t=_fp):
    return a.ctypes.data_as(t)


def band_partition(width, height, records, rank, world):
    """hm_band_partition: (row0, row1, first record, records owned, records trained on)."""
    out = (C.c_int * 5)()
    _check(lib.hm_band_partition(width, height, records, rank, world, out))
    return tuple(out)


def sample_schedule(rank, world, k):
    """spp sharding: RNG frame id (the reference's accumId) of rank `rank`'s k-th
    sample; ranks interleave, so the union over ranks of k = 0..K-1 is accumId 0..K*world-1."""
    return rank + k * world


def device_count():
    return lib.hm_device_count()


class Scene:
    """A loaded scene: geometry + BVH + lights + integrator settings (host side)."""

    def __init__(self, handle, keep=None):
        self._h = C.c_void_p(handle)
        self._keep = keep

    @classmethod
    def load(cls, config_json_path):
        h = C.c_void_p()
        _check(lib.hm_scene_load(os.fsencode(config_json_path), C.byref(h)))
        return cls(h.value)

    @classmethod
    def from_arrays(cls, *, control_points=None, segment_first_cp=None, num_strands=0, tri_vertices=None,
                    tri_normals=None, surface_kd=(0, 0, 0), surface_alpha=1.0, cam_from, cam_to=(0, 0, 0),
                    cam_up=(0, 0, 1), cos_fovy=0.66, sigma_a=(0.06, 0.1, 0.2), beta_m=0.3, beta_n=0.3,
                    alpha_deg=2.0, gains=(1, 1, 1, 1), env_rgba=None, env_scale=1.0, env_rotation=0.0,
                    dl_from=(), dl_emit=(), width, height, spp=1, path_v1=1, path_v2=40, mis=True, env_pdf=True,
                    tcnn_config_path=None):
        d = SceneDesc()
        keep = []
        if control_points is not None and len(control_points):
            cps = _f32(control_points).reshape(-1, 4)
            seg = np.ascontiguousarray(segment_first_cp, dtype=np.int32)
            keep += [cps, seg]
            d.control_points = _ptr(cps); d.num_control_points = cps.shape[0]
            d.segment_first_cp = _ptr(seg, _ip); d.num_segments = seg.shape[0]
            d.num_strands = int(num_strands)
            # real points only (phantoms excluded is what the reference does; including them
            # changes the scale slightly, so callers that care pass real bounds via hair_bounds)
            lo = np.minimum(cps[:, :3].min(axis=0), 0.0); hi = np.maximum(cps[:, :3].max(axis=0), 0.0)
            d.hair_min = (C.c_float * 3)(*lo); d.hair_max = (C.c_float * 3)(*hi)
        if tri_vertices is not None and len(tri_vertices):
            tv = _f32(tri_vertices).reshape(-1, 3); tn = _f32(tri_normals).reshape(-1, 3)
            assert tv.shape == tn.shape and tv.shape[0] % 3 == 0
            keep += [tv, tn]
            d.tri_vertices = _ptr(tv); d.tri_normals = _ptr(tn); d.num_triangles = tv.shape[0] // 3
        d.surface_kd = (C.c_float * 3)(*surface_kd); d.surface_alpha = surface_alpha
        d.cam_from = (C.c_float * 3)(*cam_from); d.cam_to = (C.c_float * 3)(*cam_to); d.cam_up = (C.c_float * 3)(*cam_up)
        d.cos_fovy = cos_fovy
        d.sigma_a = (C.c_float * 3)(*sigma_a); d.beta_m = beta_m; d.beta_n = beta_n
        d.alpha = float(np.float32(3.14159) * np.float32(alpha_deg) / np.float32(180.0))   # scene.cpp:207
        d.gains = (C.c_float * 4)(*gains)
        if env_rgba is not None:
            env = _f32(env_rgba)
            assert env.ndim == 3 and env.shape[2] == 4
            keep.append(env)
            d.env_rgba = _ptr(env); d.env_h, d.env_w = env.shape[0], env.shape[1]
        d.env_scale = env_scale; d.env_rotation = env_rotation
        dlf = _f32(dl_from).reshape(-1, 3); dle = _f32(dl_emit).reshape(-1, 3)
        keep += [dlf, dle]
        if dlf.shape[0]:
            d.dl_from = _ptr(dlf); d.dl_emit = _ptr(dle)
        d.num_dlights = dlf.shape[0]
        d.width, d.height, d.spp, d.path_v1, d.path_v2 = width, height, spp, path_v1, path_v2
        d.mis = int(mis); d.env_pdf = int(env_pdf)
        d.tcnn_config_path = os.fsencode(tcnn_config_path) if tcnn_config_path else None
        h = C.c_void_p()
        _check(lib.hm_scene_create(C.byref(d), C.byref(h)))
        return cls(h.value)

    def save_bvh_cache(self, directory):
        _check(lib.hm_scene_save_bvh_cache(self._h, os.fsencode(directory)))

    def info(self):
        i = SceneInfo()
        _check(lib.hm_scene_get_info(self._h, C.byref(i)))
        return i

    def arrays(self):
        """Host views of BVH + geometry (numpy, no copy; valid while the scene lives)."""
        i = self.info()
        nodes, cps, tv, tn, ld = _fp(), _fp(), _fp(), _fp(), _fp()
        lc, lp, sc = _ip(), _ip(), _ip()
        _check(lib.hm_scene_get_arrays(self._h, C.byref(nodes), C.byref(lc), C.byref(lp), C.byref(cps), C.byref(tv),
                                       C.byref(tn), C.byref(sc), C.byref(ld)))
        nprim = i.num_segments + i.num_triangles
        as_np = np.ctypeslib.as_array
        out = {
            "nodes": as_np(nodes, (i.num_bvh_nodes * 16,)),
            "leaf_code": as_np(lc, (nprim,)), "leaf_prim": as_np(lp, (nprim,)), "leaf_data": as_np(ld, (nprim * 16,)),
            "cps": as_np(cps, (max(i.num_control_points, 1) * 4,)) if i.num_control_points else np.zeros(4, np.float32),
            "tri_verts": as_np(tv, (i.num_triangles * 12,)) if i.num_triangles else np.zeros(12, np.float32),
            "tri_normals": as_np(tn, (i.num_triangles * 12,)) if i.num_triangles else np.zeros(12, np.float32),
            "seg_cp": as_np(sc, (i.num_segments,)) if i.num_segments else np.zeros(1, np.int32),
        }
        return out

    def env_tables(self):
        i = self.info()
        env, cpdf, ccdf, mpdf, mcdf = _fp(), _fp(), _fp(), _fp(), _fp()
        _check(lib.hm_scene_get_env_tables(self._h, C.byref(env), C.byref(cpdf), C.byref(ccdf), C.byref(mpdf), C.byref(mcdf)))
        W, H = i.env_w, i.env_h
        as_np = np.ctypeslib.as_array
        return {"env": as_np(env, (H, W, 4)), "cpdf": as_np(cpdf, (H, W + 1)), "ccdf": as_np(ccdf, (H, W + 1)),
                "mpdf": as_np(mpdf, (H + 1,)), "mcdf": as_np(mcdf, (H + 1,))}

    def close(self):
        if self._h:
            lib.hm_scene_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def bsdf_eval(wo_local, wi_local, h, sigma_a=(0.06, 0.1, 0.2), beta_m=0.3, beta_n=0.3, alpha_rad=0.0349065, gains=(1, 1, 1, 1), device=0):
    """disney_hair on the device for HOST arrays (hm_bsdf_eval): returns (f*cos [n][3], pdf [n])."""
    wo, wi, h = _f32(wo_local).reshape(-1, 3), _f32(wi_local).reshape(-1, 3), _f32(h).reshape(-1)
    n = wo.shape[0]
    s, g = _f32(sigma_a), _f32(gains)
    f = np.empty((n, 3), np.float32); pdf = np.empty(n, np.float32)
    _check(lib.hm_bsdf_eval(device, _ptr(s), C.c_float(beta_m), C.c_float(beta_n), C.c_float(alpha_rad), _ptr(g), _ptr(wo), _ptr(wi), _ptr(h), n,
                            _ptr(f), _ptr(pdf)))
    return f, pdf


def bsdf_sample(wo_local, h, rand4, sigma_a=(0.06, 0.1, 0.2), beta_m=0.3, beta_n=0.3, alpha_rad=0.0349065, gains=(1, 1, 1, 1), device=0):
    """sample_disney_hair on the device (hm_bsdf_sample): returns (wi_local [n][3], f*cos [n][3], pdf [n])."""
    wo, h, u = _f32(wo_local).reshape(-1, 3), _f32(h).reshape(-1), _f32(rand4).reshape(-1, 4)
    n = wo.shape[0]
    s, g = _f32(sigma_a), _f32(gains)
    wi = np.empty((n, 3), np.float32); f = np.empty((n, 3), np.float32); pdf = np.empty(n, np.float32)
    _check(lib.hm_bsdf_sample(device, _ptr(s), C.c_float(beta_m), C.c_float(beta_n), C.c_float(alpha_rad), _ptr(g), _ptr(wo), _ptr(h), _ptr(u), n,
                              _ptr(wi), _ptr(f), _ptr(pdf)))
    return wi, f, pdf


def load_hair_file(path):
    """The .hair reader alone (hm_hair_file_load): dict(cps [n][4], seg_cp [m], strands, bounds (min, max))."""
    counts = (C.c_int * 3)()
    _check(lib.hm_hair_file_load(os.fsencode(path), counts, None, None, None))
    cps = np.empty((counts[0], 4), np.float32); seg = np.empty(counts[1], np.int32); b = np.empty(6, np.float32)
    _check(lib.hm_hair_file_load(os.fsencode(path), counts, _ptr(cps), _ptr(seg, _ip), _ptr(b)))
    return {"cps": cps, "seg_cp": seg, "strands": counts[2], "bounds": (b[:3].copy(), b[3:].copy())}


def load_exr(path):
    """RGBA32F image [h][w][4] through the library's OpenEXR reader (hm_image_load_exr)."""
    w, h = C.c_int(), C.c_int()
    _check(lib.hm_image_load_exr(os.fsencode(path), None, C.c_size_t(0), C.byref(w), C.byref(h)))
    out = np.empty((h.value, w.value, 4), np.float32)
    _check(lib.hm_image_load_exr(os.fsencode(path), _ptr(out), C.c_size_t(out.size), C.byref(w), C.byref(h)))
    return out


def nrc_layout(width, height):
    """(numTrainingPixels, everyNth, nnFrameSize, numTrainingRecords) of render_nrc for a frame size."""
    out = (C.c_int * 4)()
    _check(lib.hm_nrc_layout(width, height, out))
    return tuple(out)


class Comm:
    """NCCL communicator of a multi-GPU job (hm_comm_*): one process per GPU.  `unique_id()` on rank 0,
    the 128 bytes travel to the other ranks by any side channel (bench.py: torchrun's TCP store)."""

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(128)
        _check(lib.hm_comm_get_unique_id(buf))
        return buf.raw

    def __init__(self, id_bytes, rank, world, device):
        assert len(id_bytes) == 128
        h = C.c_void_p()
        _check(lib.hm_comm_create(C.c_char_p(id_bytes), rank, world, device, C.byref(h)))
        self._h = h
        self.rank, self.world, self.device = rank, world, device

    def barrier(self):
        _check(lib.hm_comm_barrier(self._h))

    def all_reduce_max(self, v):
        d = C.c_double(v)
        _check(lib.hm_comm_all_reduce_max(self._h, C.byref(d)))
        return d.value

    def all_reduce_sum(self, v):
        d = C.c_double(v)
        _check(lib.hm_comm_all_reduce_sum(self._h, C.byref(d)))
        return d.value

    def close(self):
        if self._h:
            lib.hm_comm_destroy(self._h)
            self._h = None


class Mlp:
    """TINY_MLP stand-in.  `inference` / `train_step` take HOST numpy arrays; the
    `*_device` variants take raw device pointers (ints), e.g. torch tensors' data_ptr()."""

    def __init__(self, handle, owned=True):
        self._h = C.c_void_p(handle)
        self._owned = owned

    @classmethod
    def create(cls, config_path=None, in_ch=12, out_ch=3, device=0):
        h = C.c_void_p()
        _check(lib.hm_mlp_create(os.fsencode(config_path) if config_path else None, in_ch, out_ch, device, C.byref(h)))
        m = cls(h.value)
        m.in_ch, m.out_ch = in_ch, out_ch
        return m

    in_ch, out_ch = 12, 3

    @property
    def n_params(self):
        return lib.hm_mlp_n_params(self._h)

    @property
    def stream(self):
        return lib.hm_mlp_stream(self._h)

    @property
    def launch_count(self):
        return lib.hm_mlp_launch_count(self._h)

    def inference(self, x):
        x = _f32(x)
        n = x.shape[0]
        out = np.empty((n, self.out_ch), np.float32)
        _check(lib.hm_mlp_inference_host(self._h, _ptr(x), _ptr(out), n))
        return out

    def inference_device(self, d_in, d_out, n):
        _check(lib.hm_mlp_inference(self._h, C.c_void_p(d_in), C.c_void_p(d_out), n))

    def train_step(self, x, y):
        x, y = _f32(x), _f32(y)
        loss = C.c_float()
        _check(lib.hm_mlp_train_step_host(self._h, _ptr(x), _ptr(y), x.shape[0], C.byref(loss)))
        return loss.value

    def train_step_device(self, d_in, d_target, n, want_loss=False):
        loss = C.c_float()
        _check(lib.hm_mlp_train_step(self._h, C.c_void_p(d_in), C.c_void_p(d_target), n, C.byref(loss) if want_loss else None))
        return loss.value

    def forward_backward_device(self, d_in, d_target, n, n_total=0):
        _check(lib.hm_mlp_forward_backward(self._h, C.c_void_p(d_in), C.c_void_p(d_target), n, n_total))

    def gradients_device(self):
        p = C.c_void_p(); n = C.c_size_t()
        _check(lib.hm_mlp_gradients(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def optimizer_step(self):
        _check(lib.hm_mlp_optimizer_step(self._h))

    def loss(self):
        v = C.c_float()
        _check(lib.hm_mlp_loss(self._h, C.byref(v)))
        return v.value

    def get_params(self):
        out = np.empty(self.n_params, np.float32)
        _check(lib.hm_mlp_get_params(self._h, _ptr(out), C.c_size_t(out.size)))
        return out

    def set_params(self, p):
        p = _f32(p)
        _check(lib.hm_mlp_set_params(self._h, _ptr(p), C.c_size_t(p.size)))

    def reset(self):
        _check(lib.hm_mlp_reset(self._h))

    def reinitialize(self):
        _check(lib.hm_mlp_reinitialize(self._h))

    def save(self, path):
        _check(lib.hm_mlp_save(self._h, os.fsencode(path)))

    def load(self, path):
        _check(lib.hm_mlp_load(self._h, os.fsencode(path)))

    def save_snapshot(self, path):
        """tiny-cuda-nn Trainer::serialize as text JSON (what TINY_MLP::loadWeights reads)."""
        _check(lib.hm_mlp_save_snapshot(self._h, os.fsencode(path)))

    def close(self):
        if self._h and self._owned:
            lib.hm_mlp_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Renderer:
    def __init__(self, scene, kind, beta_cli=1, device=0, rank=0, world=1):
        h = C.c_void_p()
        _check(lib.hm_renderer_create(scene._h, kind, beta_cli, device, rank, world, C.byref(h)))
        self._h = h
        self.scene = scene
        self.kind = kind
        i = scene.info()
        self.W, self.H = i.width, i.height

    def render_frames(self, n=1):
        _check(lib.hm_render_frames(self._h, n))

    def render_frames_async(self, n=1):
        _check(lib.hm_render_frames_async(self._h, n))

    def flush(self):
        """Enqueue whatever render_frames_async() held back (merged tail pieces), without waiting."""
        _check(lib.hm_render_flush(self._h))

    def sync(self):
        _check(lib.hm_renderer_sync(self._h))

    def set_hair_params(self, sigma_a, beta_m, beta_n, alpha_rad, gains=(1, 1, 1, 1)):
        s, g = _f32(sigma_a), _f32(gains)
        _check(lib.hm_renderer_set_hair_params(self._h, _ptr(s), C.c_float(beta_m), C.c_float(beta_n), C.c_float(alpha_rad), _ptr(g)))

    def set_environment(self, scale, rotation=0.0):
        _check(lib.hm_renderer_set_environment(self._h, C.c_float(scale), C.c_float(rotation)))

    def set_sampling(self, mis=True, env_pdf=True):
        _check(lib.hm_renderer_set_sampling(self._h, int(mis), int(env_pdf)))

    def reset_accumulation(self):
        _check(lib.hm_renderer_reset_accumulation(self._h))

    @property
    def accum_id(self):
        return lib.hm_renderer_accum_id(self._h)

    @property
    def stream(self):
        return lib.hm_renderer_stream(self._h)

    def set_profiling(self, on):
        _check(lib.hm_renderer_set_profiling(self._h, int(on)))

    def set_profiling_stages(self, mask):
        _check(lib.hm_renderer_set_profiling_stages(self._h, C.c_uint(mask)))

    def set_profiling_period(self, n):
        _check(lib.hm_renderer_set_profiling_period(self._h, int(n)))

    def set_collect_stats(self, on):
        _check(lib.hm_renderer_set_collect_stats(self._h, int(on)))

    def set_skip_unused_queries(self, on):
        _check(lib.hm_renderer_set_skip_unused_queries(self._h, int(on)))

    def reset_stats(self):
        _check(lib.hm_renderer_reset_stats(self._h))

    def set_comm(self, comm):
        """Attach a Comm: sample schedule by group, gradient all-reduce inside every training step,
        reduce_framebuffers() (hm_renderer_set_comm)."""
        _check(lib.hm_renderer_set_comm(self._h, comm._h if comm is not None else None))
        self._comm = comm

    def reduce_framebuffers(self):
        _check(lib.hm_reduce_framebuffers(self._h))

    def set_frame_schedule(self, offset, stride):
        _check(lib.hm_renderer_set_frame_schedule(self._h, offset, stride))

    def stats(self):
        s = Stats()
        _check(lib.hm_renderer_get_stats(self._h, C.byref(s)))
        return s

    def mlp(self):
        h = lib.hm_renderer_mlp(self._h)
        if not h:
            raise HairMSNNError(-4, "this renderer kind has no MLP")
        m = Mlp(h, owned=False)
        m.in_ch = self.layout()[0]
        return m

    def msnn_trace(self): _check(lib.hm_msnn_trace(self._h))
    def msnn_train_backward(self): _check(lib.hm_msnn_train_backward(self._h))
    def msnn_train_apply(self): _check(lib.hm_msnn_train_apply(self._h))
    def msnn_finish(self): _check(lib.hm_msnn_finish(self._h))
    def msnn_pretrain(self, steps): _check(lib.hm_msnn_pretrain(self._h, steps))
    def msnn_train_data_gen(self): _check(lib.hm_msnn_train_data_gen(self._h))

    # render_nrc split frame (render_nrc.cu:640-700)
    def nrc_trace(self): _check(lib.hm_nrc_trace(self._h))
    def nrc_query(self): _check(lib.hm_nrc_query(self._h))
    def nrc_train_backward(self): _check(lib.hm_nrc_train_backward(self._h))
    def nrc_train_apply(self): _check(lib.hm_nrc_train_apply(self._h))
    def nrc_end(self): _check(lib.hm_nrc_end(self._h))
    def nrc_set_all_unbiased(self, on): _check(lib.hm_nrc_set_all_unbiased(self._h, int(on)))

    def layout(self):
        """(MLP input channels, inference rows per frame, training records per step, everyNth)"""
        out = (C.c_int * 4)()
        _check(lib.hm_renderer_get_layout(self._h, out))
        return tuple(out)

    def nrc_train_records(self):
        """TrainBuffer contents after nrc_trace(): dict of arrays [pixels][40][3] + bounces/hit [pixels]."""
        _, nbytes = self.device_buffer(BUF_NRC_TRAIN_RECORDS)
        raw = np.empty(nbytes, np.uint8)
        _check(lib.hm_get_buffer(self._h, BUF_NRC_TRAIN_RECORDS, raw.ctypes.data_as(C.c_void_p), C.c_size_t(nbytes)))
        rec = raw.view(np.float32).reshape(-1, 5 * NRC_MAX_BOUNCES * 3 + 2)
        body = rec[:, :-2].reshape(-1, 5, NRC_MAX_BOUNCES, 3)
        tail = rec[:, -2:].copy().view(np.int32)
        return {"vert": body[:, 0], "wo": body[:, 1], "n": body[:, 2], "radiance": body[:, 3], "beta": body[:, 4],
                "bounces": tail[:, 0], "hit": tail[:, 1]}

    def device_buffer(self, which):
        p = C.c_void_p(); n = C.c_size_t()
        _check(lib.hm_get_device_buffer(self._h, which, C.byref(p), C.byref(n)))
        return p.value, n.value

    def buffer(self, which):
        _, nbytes = self.device_buffer(which)
        if which in (BUF_FB8,):
            out = np.empty((self.H, self.W), np.uint32)
        elif which in (BUF_TRAIN_IDXS, BUF_SCENE_INDICES):
            out = np.empty(nbytes // 4, np.int32)
        elif which in (BUF_NN_FRAME_INPUT, BUF_NN_FRAME_OUTPUT, BUF_NN_TRAIN_INPUT, BUF_NN_TRAIN_OUTPUT, BUF_SCENE_POINTS,
                       BUF_ENV_CPDF, BUF_ENV_CCDF, BUF_ENV_MPDF, BUF_ENV_MCDF):
            out = np.empty(nbytes // 4, np.float32)
        elif which == BUF_GBUFFER:
            out = np.empty((nbytes // (16 * self.W), self.W, 4), np.float32)   # this renderer's row band
        else:
            out = np.empty((self.H, self.W, 4), np.float32)
        _check(lib.hm_get_buffer(self._h, which, out.ctypes.data_as(C.c_void_p), C.c_size_t(out.nbytes)))
        return out

    def readback_async(self, which, host_ptr, nbytes):
        _check(lib.hm_readback_async(self._h, which, C.c_void_p(host_ptr), C.c_size_t(nbytes)))

    def readback_rows_async(self, which, row0, rows, host_ptr):
        _check(lib.hm_readback_rows_async(self._h, which, row0, rows, C.c_void_p(host_ptr)))

    def rows(self):
        out = (C.c_int * 2)()
        _check(lib.hm_renderer_get_rows(self._h, out))
        return out[0], out[1]

    def trace_rays(self, org, dir, any_hit=False, tmin=0.0, tmax=1e30, stats=False):
        org, dir = _f32(org).reshape(-1, 3), _f32(dir).reshape(-1, 3)
        n = org.shape[0]
        hit = np.empty((n, 4), np.float32)
        st = np.empty((n, 2), np.int32) if stats else None
        _check(lib.hm_trace_rays(self._h, _ptr(org), _ptr(dir), n, int(any_hit), C.c_float(tmin), C.c_float(tmax),
                                 _ptr(hit), _ptr(st, _ip) if stats else None))
        res = {"t": hit[:, 0].copy(), "prim": hit[:, 1].copy().view(np.int32), "u": hit[:, 2].copy(), "v": hit[:, 3].copy()}
        if stats:
            res["nodes"], res["prims"] = st[:, 0], st[:, 1]
        return res

    def trace_rays_device(self, d_org, d_dir, n, d_out, any_hit=False, tmin=0.0, tmax=1e30):
        _check(lib.hm_trace_rays_device(self._h, C.c_void_p(d_org), C.c_void_p(d_dir), n, int(any_hit), C.c_float(tmin),
                                        C.c_float(tmax), C.c_void_p(d_out)))

    def save_png(self, path): _check(lib.hm_save_png(self._h, os.fsencode(path)))
    def save_exr(self, path, which=BUF_FINAL_AVG): _check(lib.hm_save_exr(self._h, which, os.fsencode(path)))
    def write_stats(self, path): _check(lib.hm_write_stats(self._h, os.fsencode(path)))

    def close(self):
        if self._h:
            lib.hm_renderer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
