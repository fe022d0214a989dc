// ref_scene.cpp — TEST INFRASTRUCTURE (oracle/_ref).  The reference's own scene.cpp compiled for the host from where
// it lies (nothing copied): exposes Scene::extractHairData (scene.cpp:10-73) over cyHairFile::LoadFromFile so that the
// product's .hair reader (hm_io.cpp: load_hair_file) can be compared array by array.
#include "scene.cpp"

// parseScene (same translation unit) calls into model.cpp, which is not part of this build; the hooks below never do
Model* loadOBJ(const std::string&) { throw std::runtime_error("ref_scene: loadOBJ is not built"); }
bool loadEnvTexture(std::string&, Texture*) { throw std::runtime_error("ref_scene: loadEnvTexture is not built"); }

static Scene g_scene;

extern "C" {

// Loads a .hair file with the reference's loader and runs extractHairData.  counts[3] = control points, segments, strands.
int ref_extract_hair(const char* path, int* counts) {
    g_scene.hair = cyHairFile();
    g_scene.hairModel = HairModel();
    if (g_scene.hair.LoadFromFile(path) < 0) return -1;
    g_scene.extractHairData();
    counts[0] = (int)g_scene.hairModel.controlPoints.size();
    counts[1] = (int)g_scene.hairModel.segmentIndices.size();
    counts[2] = g_scene.hairModel.numStrands;
    return 0;
}
// control points [n][3], widths [n] (already 0.2 x file thickness), segment indices [m]; bounds6 = min, max; scale
void ref_hair_arrays(float* cps3, float* widths, int* seg_idx, float* bounds6, float* scale) {
    const HairModel& h = g_scene.hairModel;
    for (size_t i = 0; i < h.controlPoints.size(); ++i) {
        cps3[3 * i] = h.controlPoints[i].x; cps3[3 * i + 1] = h.controlPoints[i].y; cps3[3 * i + 2] = h.controlPoints[i].z;
        widths[i] = h.widths[i];
    }
    for (size_t i = 0; i < h.segmentIndices.size(); ++i) seg_idx[i] = h.segmentIndices[i];
    bounds6[0] = h.minBound.x; bounds6[1] = h.minBound.y; bounds6[2] = h.minBound.z;
    bounds6[3] = h.maxBound.x; bounds6[4] = h.maxBound.y; bounds6[5] = h.maxBound.z;
    *scale = h.scale;
}

}  // extern "C"
