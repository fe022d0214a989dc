"""TEST INFRASTRUCTURE: ctypes binding of oracle/_ref/libref_{pt,msnn}.so — the
REFERENCE's own per-path sources compiled for the host (oracle/Makefile).  Only tests,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_up = C.POINTER(C.c_uint32)


def ensure_built():
    """Builds oracle/_ref when the reference tree is present; otherwise the prebuilt
    libraries shipped with the snapshot must exist."""
    need = [os.path.join(REF_DIR, f"libref_{k}.so") for k in ("pt", "msnn", "nrc")]
    if os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    for p in need:
        if not os.path.exists(p):
            raise FileNotFoundError(f"{p} missing (build it where /root/reference exists: make -C oracle)")


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t=_fp):
    return a.ctypes.data_as(t)


class RefHost:
    """One loaded copy of the reference host build with a scene bound to it."""

    def __init__(self, which="pt"):
        ensure_built()
        self.lib = C.CDLL(os.path.join(REF_DIR, f"libref_{which}.so"))
        self.which = which
        self._keep = []
        self.lib.ref_rng_seed.restype = C.c_uint32

    # -- scene plumbing ---------------------------------------------------------------
    def bind_scene(self, scene, mis=True, path_v1=1, path_v2=40):
        """scene: hairmsnn_b200.api.Scene (host arrays are shared, not copied)."""
        L = self.lib
        info = scene.info()
        arr = scene.arrays()
        self._keep.append(arr)
        ns, nt = info.num_segments, info.num_triangles
        L.ref_set_geometry(_p(arr["nodes"]), info.num_bvh_nodes, _p(arr["leaf_code"], _ip), _p(arr["leaf_prim"], _ip),
                           _p(arr["cps"]), _p(arr["tri_verts"]), _p(arr["seg_cp"], _ip), ns, nt, _p(arr["leaf_data"]))
        # reference triangle layout: flattened float3 soup + index triples
        tv = np.ascontiguousarray(arr["tri_verts"].reshape(-1, 4)[:, :3]) if nt else np.zeros((3, 3), np.float32)
        tn = np.ascontiguousarray(arr["tri_normals"].reshape(-1, 4)[:, :3]) if nt else np.zeros((3, 3), np.float32)
        idx = np.arange(max(3 * nt, 3), dtype=np.int32)
        uv = np.zeros((max(3 * nt, 3), 2), np.float32)
        kd = np.zeros(3, np.float32)
        self._keep += [tv, tn, idx, uv, kd]
        L.ref_set_triangle_mesh(_p(tv), _p(tn), _p(idx, _ip), _p(uv), _p(kd), C.c_float(1.0))
        L.ref_clear_textures()
        has_env = info.env_w > 0
        ids = [0] * 5
        if has_env:
            t = scene.env_tables()
            self._keep.append(t)
            W, H = info.env_w, info.env_h
            ids[0] = L.ref_add_texture(_p(t["env"]), W, H, 4, 1)
            ids[1] = L.ref_add_texture(_p(t["cpdf"]), W + 1, H, 1, 0)
            ids[2] = L.ref_add_texture(_p(t["ccdf"]), W + 1, H, 1, 0)
            ids[3] = L.ref_add_texture(_p(t["mpdf"]), H + 1, 1, 1, 0)
            ids[4] = L.ref_add_texture(_p(t["mcdf"]), H + 1, 1, 1, 0)
        self.scene_kw = getattr(scene, "kw", None)
        return ids, has_env

    def set_lights(self, ids, has_env, env_pdf, env_scale, env_rot, env_w, env_h, dl_from, dl_emit):
        dl_from = _f(dl_from).reshape(-1, 3)
        if dl_from.shape[0]:
            dl_from = dl_from / np.sqrt((dl_from.astype(np.float32) ** 2).sum(axis=1, keepdims=True, dtype=np.float32))
        dl_from = _f(dl_from); dl_emit = _f(dl_emit).reshape(-1, 3)
        self._keep += [dl_from, dl_emit]
        self.lib.ref_set_lights(int(has_env), int(env_pdf), C.c_float(env_scale), C.c_float(env_rot), env_w, env_h,
                                ids[0], ids[1], ids[2], ids[3], ids[4], dl_from.shape[0], _p(dl_from), _p(dl_emit))

    def set_camera(self, info):
        self.lib.ref_set_camera(info.cam_pos, info.cam_d00, info.cam_du, info.cam_dv)

    def set_hair(self, sigma_a, beta_m, beta_n, alpha_rad, gains=(1, 1, 1, 1)):
        s = _f(sigma_a); g = _f(gains)
        self.lib.ref_set_hair(_p(s), C.c_float(beta_m), C.c_float(beta_n), C.c_float(alpha_rad), _p(g))

    def set_integrator(self, mis, path_v1, path_v2, scene_scale):
        self.lib.ref_set_integrator(int(mis), path_v1, path_v2, C.c_float(scene_scale))

    def bind_all(self, scene, kw):
        """kw: the keyword dict the scene was created from (api.Scene.from_arrays)."""
        info = scene.info()
        ids, has_env = self.bind_scene(scene)
        self.set_lights(ids, has_env, kw.get("env_pdf", True), kw.get("env_scale", 1.0), kw.get("env_rotation", 0.0),
                        info.env_w, info.env_h, kw.get("dl_from", ()), kw.get("dl_emit", ()))
        self.set_camera(info)
        alpha = float(np.float32(3.14159) * np.float32(kw.get("alpha_deg", 2.0)) / np.float32(180.0))
        self.set_hair(kw.get("sigma_a", (0.06, 0.1, 0.2)), kw.get("beta_m", 0.3), kw.get("beta_n", 0.3), alpha,
                      kw.get("gains", (1, 1, 1, 1)))
        self.set_integrator(kw.get("mis", True), kw.get("path_v1", 1), kw.get("path_v2", 40), info.scene_scale)
        return info

    # -- entry points -----------------------------------------------------------------
    def render_pt(self, accum_id, W, H, accum=None, average=None, x0=0, y0=0, x1=None, y1=None, threads=0):
        assert self.which == "pt"
        x1 = W if x1 is None else x1; y1 = H if y1 is None else y1
        accum = np.zeros((H, W, 4), np.float32) if accum is None else accum
        average = np.zeros((H, W, 4), np.float32) if average is None else average
        fb = np.zeros((H, W), np.uint32)
        self.lib.ref_render_pt(accum_id, x0, y0, x1, y1, W, H, _p(accum), _p(average), _p(fb, _up), threads or os.cpu_count())
        return accum, average, fb

    def render_msnn_gbuffer(self, accum_id, W, H, beta, every_nth, train_idxs, in_ch=12, y0=0, y1=None, threads=0):
        assert self.which == "msnn"
        y1 = H if y1 is None else y1
        train_idxs = np.ascontiguousarray(train_idxs, dtype=np.int32)
        rec = train_idxs.shape[0]
        nn_in = np.zeros((W * H, in_ch), np.float32)
        tr_in = np.zeros((rec, in_ch), np.float32); tr_out = np.zeros((rec, 3), np.float32)
        gb = np.zeros((W * H, 8), np.float32)
        self.lib.ref_render_msnn_gbuffer(accum_id, y0, y1, W, H, beta, every_nth, _p(train_idxs, _ip), in_ch, _p(nn_in),
                                         _p(tr_in), _p(tr_out), _p(gb), threads or os.cpu_count())
        return nn_in, tr_in, tr_out, gb

    def render_msnn_rows(self, accum_id, W, H, beta, every_nth, train_idxs, rows, bufs=None, in_ch=12, threads=0):
        """G_BUFFER pass for a list of rows, all host threads sharing the rows' pixels.  `bufs` = (nn_in [W*H][in_ch],
        tr_in, tr_out, gb [len(rows)*W][8]) from a previous call are reused."""
        assert self.which == "msnn"
        train_idxs = np.ascontiguousarray(train_idxs, dtype=np.int32)
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        rec = train_idxs.shape[0]
        if bufs is None:
            bufs = (np.zeros((W * H, in_ch), np.float32), np.zeros((rec, in_ch), np.float32), np.zeros((rec, 3), np.float32),
                    np.zeros((len(rows) * W, 8), np.float32))
        nn_in, tr_in, tr_out, gb = bufs
        self.lib.ref_render_msnn_gbuffer_rows(accum_id, _p(rows, _ip), len(rows), W, H, beta, every_nth, _p(train_idxs, _ip), in_ch,
                                              _p(nn_in), _p(tr_in), _p(tr_out), _p(gb), threads or os.cpu_count())
        return bufs

    def render_msnn_composite(self, accum_id, W, H, hit, is_surface, short_color, nn_out, pt_accum, nn_accum, final_accum):
        """RENDER pass of cuda/hair_msnn.cu:314-356 over all pixels.  hit / is_surface: bool [H*W]; short_color [H*W][3];
        nn_out [H*W][3]; the three accumulation buffers float4 [H][W][4] are updated in place.
        Returns (pt_avg, nn_avg, final_avg, fb8)."""
        assert self.which == "msnn"
        gb = np.zeros((W * H, 8), np.float32)
        gb[:, 0] = np.asarray(hit, np.float32).reshape(-1); gb[:, 1] = np.asarray(is_surface, np.float32).reshape(-1)
        gb[:, 5:8] = _f(short_color).reshape(-1, 3)
        nn = _f(nn_out).reshape(-1, 3)
        pt_avg = np.zeros((H, W, 4), np.float32); nn_avg = np.zeros((H, W, 4), np.float32); final_avg = np.zeros((H, W, 4), np.float32)
        fb = np.zeros((H, W), np.uint32)
        self.lib.ref_render_msnn_composite(accum_id, W, H, _p(gb), _p(nn), _p(pt_accum), _p(nn_accum), _p(final_accum),
                                           _p(pt_avg), _p(nn_avg), _p(final_avg), _p(fb, _up))
        return pt_avg, nn_avg, final_avg, fb

    def render_msnn_train_data_gen(self, accum_id, W, H, beta, scene_indices, sampled_points, in_ch=12, threads=0):
        assert self.which == "msnn"
        idx = np.ascontiguousarray(scene_indices, dtype=np.int32)
        pts = _f(sampled_points).reshape(-1, 3)
        tr_in = np.zeros((16384, in_ch), np.float32); tr_out = np.zeros((16384, 3), np.float32)
        self.lib.ref_render_msnn_train_data_gen(accum_id, W, H, beta, _p(idx, _ip), _p(pts), in_ch, _p(tr_in), _p(tr_out),
                                                threads or os.cpu_count())
        return tr_in, tr_out

    def render_nrc_gbuffer(self, accum_id, W, H, every_nth, train_idxs, nn_frame_rows, c=0.01, all_unbiased=False,
                           in_ch=9, threads=0):
        """G_BUFFER pass of cuda/nrc.cu.  Returns (nn_frame_in [rows][in_ch], gbuffer [W*H][8] = hit,
        pathRadiance, beta, bounces).  The path records stay inside the library for render_nrc_render /
        nrc_train_record."""
        assert self.which == "nrc"
        ntp = len(train_idxs)
        # the reference reads trainIdxs[ntp] when W*H is not a multiple of everyNth; give that
        # group an index no pixel of it can match
        idx = np.concatenate([np.asarray(train_idxs, np.int32), np.array([every_nth - 1], np.int32)])
        assert W * H - ntp * every_nth < every_nth
        self._nrc_idx = np.ascontiguousarray(idx)
        nn_in = np.zeros((nn_frame_rows, in_ch), np.float32)
        gb = np.zeros((W * H, 8), np.float32)
        self.lib.ref_render_nrc_gbuffer(accum_id, W, H, every_nth, _p(self._nrc_idx, _ip), ntp, C.c_float(c),
                                        int(all_unbiased), in_ch, _p(nn_in), _p(gb), threads or os.cpu_count())
        return nn_in, gb

    def nrc_train_record(self, tr_ofs):
        out = np.zeros(2 + 15 * 40, np.float32)
        self.lib.ref_nrc_train_record(tr_ofs, _p(out))
        body = out[2:].reshape(40, 5, 3)
        return {"bounces": int(out[0]), "hit": int(out[1]), "vert": body[:, 0], "wo": body[:, 1], "n": body[:, 2],
                "radiance": body[:, 3], "beta": body[:, 4]}

    def render_nrc_render(self, accum_id, W, H, every_nth, nn_out, records, in_ch=9, accum=None, average=None):
        """RENDER pass of cuda/nrc.cu on the state left by render_nrc_gbuffer."""
        assert self.which == "nrc"
        nn_out = _f(nn_out)
        tr_in = np.zeros((records, in_ch), np.float32); tr_gt = np.zeros((records, 3), np.float32)
        accum = np.zeros((H, W, 4), np.float32) if accum is None else accum
        average = np.zeros((H, W, 4), np.float32) if average is None else average
        fb = np.zeros((H, W), np.uint32)
        self.lib.ref_render_nrc_render(accum_id, W, H, every_nth, _p(self._nrc_idx, _ip), in_ch, _p(nn_out), _p(tr_in),
                                       _p(tr_gt), _p(accum), _p(average), _p(fb, _up))
        return tr_in, tr_gt, accum, average, fb

    def trace_ids(self, org, dir, any_hit=False):
        """(t, prim, u) of the host traversal (binary tree + the product's intersector compiled for the host)."""
        org, dir = _f(org).reshape(-1, 3), _f(dir).reshape(-1, 3)
        n = org.shape[0]
        t = np.zeros(n, np.float32); p = np.zeros(n, np.int32); u = np.zeros(n, np.float32)
        self.lib.ref_trace_ids(n, _p(org), _p(dir), int(any_hit), _p(t), _p(p, _ip), _p(u))
        return t, p, u

    def trace_radiance(self, org, dir):
        org, dir = _f(org).reshape(-1, 3), _f(dir).reshape(-1, 3)
        out = np.zeros((org.shape[0], 24), np.float32)
        self.lib.ref_trace_radiance(org.shape[0], _p(org), _p(dir), _p(out))
        return out
