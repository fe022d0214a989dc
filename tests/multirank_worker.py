"""Worker of tests/test_gpu_multirank.py: one process per GPU (rank r on device r), communicator id through a file.

    python tests/multirank_worker.py <rank> <world> <id_file> <out_dir> <mode> <frames>

mode: pt_spp | pt_bands | msnn_spp | msnn_bands | nrc_spp.  Writes <out_dir>/<mode>_rank<r>.npz.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def get_id(path, rank):
    from hairmsnn_b200 import api
    if rank == 0:
        ident = api.Comm.unique_id()
        with open(path + ".tmp", "wb") as f:
            f.write(ident)
        os.rename(path + ".tmp", path)
        return ident
    for _ in range(600):
        if os.path.exists(path):
            b = open(path, "rb").read()
            if len(b) == 128:
                return b
        time.sleep(0.1)
    raise RuntimeError("no communicator id")


def main():
    rank, world, id_file, out_dir, mode, frames = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], sys.argv[5], int(sys.argv[6])
    from hairmsnn_b200 import api
    from common import small_scene_kwargs
    kw = small_scene_kwargs(width=128, height=128, strands=800, segs=12, path_v2=8)
    sc = api.Scene.from_arrays(**kw)
    comm = api.Comm(get_id(id_file, rank), rank, world, rank)
    kind = {"pt": api.PATH_TRACING, "msnn": api.HAIR_MSNN, "nrc": api.NRC}[mode.split("_")[0]]
    bands = mode.endswith("bands")
    r = api.Renderer(sc, kind, beta_cli=1, device=rank, rank=rank if bands else 0, world=world if bands else 1)
    r.set_comm(comm)
    if kind == api.HAIR_MSNN:
        r.msnn_pretrain(3)
    r.render_frames(frames)
    out = {"local_final_accum": r.buffer(api.BUF_FINAL_ACCUM)}
    if kind != api.PATH_TRACING:
        out["params"] = r.mlp().get_params()
        out["loss"] = np.float32(r.stats().last_loss)
    r.reduce_framebuffers()
    out["final_avg"] = r.buffer(api.BUF_FINAL_AVG)
    out["fb8"] = r.buffer(api.BUF_FB8)
    if kind == api.HAIR_MSNN:
        out["pt_avg"] = r.buffer(api.BUF_PT_AVG)
        out["nn_avg"] = r.buffer(api.BUF_NN_AVG)
    comm.barrier()
    np.savez(os.path.join(out_dir, f"{mode}_rank{rank}.npz"), **out)
    r.close()
    comm.close()


if __name__ == "__main__":
    main()
