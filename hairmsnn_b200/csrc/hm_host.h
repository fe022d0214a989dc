// hm_host.h — host-side scene containers shared by the loaders, the BVH builder and
// the C-ABI implementation.
#pragma once
#include <string>
#include <vector>

#include "hm_bvh.h"

namespace hm {

// Fibre + head geometry in the layout the kernels consume.
struct HostGeometry {
    std::vector<F4> cps;          // Catmull-Rom control points incl. phantom endpoints; w = radius
    std::vector<int> seg_cp;      // per segment: index of the first of its 4 control points
    std::vector<int> seg_strand;  // per segment: strand id
    std::vector<F4> tri_verts;    // flattened triangle soup, 3 per triangle
    std::vector<F4> tri_normals;  // per-corner normals, 3 per triangle
    std::vector<float> tri_uv;    // per-corner texcoords, 6 per triangle
    int num_strands = 0;
    // bounds exactly as the reference accumulates them (hair bounds start at the
    // origin, headers/model.h:93-94; mesh bounds use the first corner of each
    // triangle only, model.cpp:309-310)
    float hair_min[3] = {0, 0, 0}, hair_max[3] = {0, 0, 0};
    float mesh_min[3] = {1e30f, 1e30f, 1e30f}, mesh_max[3] = {-1e30f, -1e30f, -1e30f};
    float hair_scale = 0.f;
    float scene_scale = 0.f;
    float kd[3] = {0, 0, 0};      // head diffuse colour (tinyobj default 0 when the .mtl has no Kd)
    float surf_alpha = 1.f;
};

struct HostBvh {
    std::vector<F4> nodes;
    std::vector<int> leaf_code;
    std::vector<int> leaf_prim;
};

void build_bvh(const HostGeometry& geo, HostBvh& out, int threads_hint = 0);

inline GeomView make_view(const HostGeometry& g, const HostBvh& b) {
    GeomView v;
    v.nodes = b.nodes.data();
    v.leaf_code = b.leaf_code.data();
    v.leaf_prim = b.leaf_prim.data();
    v.cps = g.cps.data();
    v.tri_verts = g.tri_verts.data();
    v.num_segments = (int)g.seg_cp.size();
    v.num_tris = (int)(g.tri_verts.size() / 3);
    v.num_nodes = (int)(b.nodes.size() / 4);
    return v;
}

}  // namespace hm
