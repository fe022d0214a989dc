// hm_io.h — on-disk formats either side of the path: config.json / tcnn JSON, Cem Yuksel
// .hair, Wavefront .obj, OpenEXR (read: NONE/ZIPS/ZIP/PIZ scanline; write: NONE), PNG.
// Load/save-time CPU plumbing written from the format
// specifications; the reference uses nlohmann/json, cyHairFile, tinyobjloader, tinyexr
// and stb for the same jobs.
#pragma once
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "hm_host.h"

namespace hm {

struct IoError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// ---- minimal JSON (RFC 8259 subset: no \u surrogate pairs) ----
struct Json {
    enum Type { Null, Bool, Number, String, Array, Object } type = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;

    const Json* find(const std::string& key) const;
    const Json& at(const std::string& key) const;   // throws std::invalid_argument
    bool has(const std::string& key) const { return find(key) != nullptr; }
    double number() const;
    bool boolean() const;
    const std::string& string() const;
};
Json parse_json(const std::string& text);
Json parse_json_file(const std::string& path);

// Resolves a path from a scene file: as written; else relative to base_dir by
// progressively shorter suffixes; file-name match is case-insensitive.
std::string resolve_scene_path(const std::string& written, const std::string& base_dir);

// parseScene (scene.cpp:119-339): fills everything except the BVH / env tables / scales.
void load_scene_file(const std::string& config_path, HostScene& out);

// Scene::extractHairData (scene.cpp:10-73) over a .hair file
void load_hair_file(const std::string& path, HostGeometry& geo);
// loadOBJ (model.cpp:233-330): triangulated, flattened soup with per-corner normals
void load_obj_file(const std::string& path, HostGeometry& geo);
// LoadEXR semantics (tinyexr): RGBA float, alpha = 1 when absent
void load_exr_rgba(const std::string& path, std::vector<float>& rgba, int& w, int& h);

// OWLViewer::screenShot (OWLViewer.cpp:109-124): rows flipped, alpha forced opaque
void write_png_flipped(const std::string& path, const uint32_t* rgba8, int w, int h);
// saveEXR (model.cpp:366-383): rows flipped, RGBA fp32
void write_exr_flipped(const std::string& path, const float* rgba, int w, int h);

}  // namespace hm
