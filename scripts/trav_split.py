"""Nodes / primitives per ray by ray kind on the bench scene (instrumented traversal)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from hairmsnn_b200 import api
sc, kw = bench.make_scene(50000)
r = api.Renderer(sc, api.HAIR_MSNN, beta_cli=1)
r.render_frames(4); r.reset_stats(); r.set_collect_stats(True); r.render_frames(4); s = r.stats()
print(f"primary: {s.trav_nodes_primary / s.rays_primary:.1f} nodes {s.trav_prims_primary / s.rays_primary:.2f} prims per ray ({s.rays_primary // 4} rays/frame)")
print(f"extend : {s.trav_nodes_extend / s.rays_extend:.1f} nodes {s.trav_prims_extend / s.rays_extend:.2f} prims per ray ({s.rays_extend // 4} rays/frame)")
print(f"shadow : {s.trav_nodes_shadow / s.rays_shadow:.1f} nodes {s.trav_prims_shadow / s.rays_shadow:.2f} prims per ray ({s.rays_shadow // 4} rays/frame)")
