// hm_io.cpp — see hm_io.h.
#include "hm_io.h"

#include <dirent.h>
#include <sys/stat.h>
#include <zlib.h>

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

namespace hm {

// =====================================================================================
// JSON
// =====================================================================================
const Json* Json::find(const std::string& key) const {
    if (type != Object) return nullptr;
    for (const auto& kv : obj)
        if (kv.first == key) return &kv.second;
    return nullptr;
}
const Json& Json::at(const std::string& key) const {
    const Json* j = find(key);
    if (!j) throw std::invalid_argument("json: missing key '" + key + "'");
    return *j;
}
double Json::number() const {
    if (type != Number) throw std::invalid_argument("json: expected a number");
    return num;
}
bool Json::boolean() const {
    if (type == Bool) return b;
    if (type == Number) return num != 0.0;
    throw std::invalid_argument("json: expected a boolean");
}
const std::string& Json::string() const {
    if (type != String) throw std::invalid_argument("json: expected a string");
    return str;
}

namespace {
struct JsonParser {
    const std::string& s;
    size_t i = 0;
    explicit JsonParser(const std::string& t) : s(t) {}
    [[noreturn]] void fail(const char* what) { throw std::invalid_argument(std::string("json: ") + what + " at offset " + std::to_string(i)); }
    void ws() { while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) ++i; }
    Json value() {
        ws();
        if (i >= s.size()) fail("unexpected end");
        char c = s[i];
        Json j;
        if (c == '{') {
            j.type = Json::Object; ++i; ws();
            if (i < s.size() && s[i] == '}') { ++i; return j; }
            for (;;) {
                ws();
                if (i >= s.size() || s[i] != '"') fail("expected key");
                std::string k = str();
                ws();
                if (i >= s.size() || s[i] != ':') fail("expected ':'");
                ++i;
                j.obj.emplace_back(k, value());
                ws();
                if (i < s.size() && s[i] == ',') { ++i; continue; }
                if (i < s.size() && s[i] == '}') { ++i; break; }
                fail("expected ',' or '}'");
            }
        } else if (c == '[') {
            j.type = Json::Array; ++i; ws();
            if (i < s.size() && s[i] == ']') { ++i; return j; }
            for (;;) {
                j.arr.push_back(value());
                ws();
                if (i < s.size() && s[i] == ',') { ++i; continue; }
                if (i < s.size() && s[i] == ']') { ++i; break; }
                fail("expected ',' or ']'");
            }
        } else if (c == '"') {
            j.type = Json::String; j.str = str();
        } else if (s.compare(i, 4, "true") == 0) { j.type = Json::Bool; j.b = true; i += 4; }
        else if (s.compare(i, 5, "false") == 0) { j.type = Json::Bool; j.b = false; i += 5; }
        else if (s.compare(i, 4, "null") == 0) { j.type = Json::Null; i += 4; }
        else {
            size_t b = i;
            while (i < s.size() && (isdigit((unsigned char)s[i]) || s[i] == '-' || s[i] == '+' || s[i] == '.' || s[i] == 'e' || s[i] == 'E')) ++i;
            if (b == i) fail("unexpected character");
            j.type = Json::Number;
            j.num = strtod(s.substr(b, i - b).c_str(), nullptr);
        }
        return j;
    }
    std::string str() {
        std::string out;
        ++i;
        while (i < s.size() && s[i] != '"') {
            char c = s[i++];
            if (c == '\\') {
                if (i >= s.size()) fail("bad escape");
                char e = s[i++];
                switch (e) {
                    case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break; case 'f': out += '\f'; break;
                    case 'u': {
                        if (i + 4 > s.size()) fail("bad \\u escape");
                        unsigned cp = (unsigned)strtoul(s.substr(i, 4).c_str(), nullptr, 16); i += 4;
                        if (cp < 0x80) out += (char)cp;
                        else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
                        else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
                        break;
                    }
                    default: out += e;
                }
            } else out += c;
        }
        if (i >= s.size()) fail("unterminated string");
        ++i;
        return out;
    }
};

std::string read_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw IoError("cannot open " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}
bool file_exists(const std::string& p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}
std::string lower(std::string s) {
    for (auto& c : s) c = (char)tolower((unsigned char)c);
    return s;
}
}  // namespace

Json parse_json(const std::string& text) {
    JsonParser p(text);
    Json j = p.value();
    p.ws();
    if (p.i != text.size()) p.fail("trailing characters");
    return j;
}
Json parse_json_file(const std::string& path) { return parse_json(read_file(path)); }

std::string resolve_scene_path(const std::string& written, const std::string& base_dir) {
    if (file_exists(written)) return written;
    std::string w = written;
    for (auto& c : w) if (c == '\\') c = '/';
    // candidate suffixes: a/b/c.ext, b/c.ext, c.ext — tried against base_dir and its parents
    std::vector<std::string> parts;
    {
        std::stringstream ss(w);
        std::string item;
        while (std::getline(ss, item, '/')) if (!item.empty()) parts.push_back(item);
    }
    auto try_ci = [](const std::string& dir, const std::string& rel) -> std::string {
        // walk rel component by component, case-insensitively
        std::string cur = dir.empty() ? "." : dir;
        std::stringstream ss(rel);
        std::string item;
        while (std::getline(ss, item, '/')) {
            if (item.empty()) continue;
            std::string exact = cur + "/" + item;
            struct stat st;
            if (stat(exact.c_str(), &st) == 0) { cur = exact; continue; }
            DIR* d = opendir(cur.c_str());
            if (!d) return "";
            std::string found;
            while (dirent* e = readdir(d))
                if (lower(e->d_name) == lower(item)) { found = e->d_name; break; }
            closedir(d);
            if (found.empty()) return "";
            cur += "/" + found;
        }
        return file_exists(cur) ? cur : "";
    };
    std::string dir = base_dir.empty() ? "." : base_dir;
    for (int up = 0; up < 3; ++up) {
        for (size_t start = 0; start < parts.size(); ++start) {
            std::string rel;
            for (size_t k = start; k < parts.size(); ++k) rel += (k > start ? "/" : "") + parts[k];
            std::string r = try_ci(dir, rel);
            if (!r.empty()) return r;
        }
        dir += "/..";
    }
    throw IoError("cannot find '" + written + "' (also tried relative to " + base_dir + ")");
}

// =====================================================================================
// .hair  (Cem Yuksel HAIR format: 128-byte header "HAIR", counts, bit flags, defaults)
// =====================================================================================
void load_hair_file(const std::string& path, HostGeometry& geo) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw IoError("cannot open hair file " + path);
    struct Header {
        char sig[4];
        uint32_t hair_count, point_count, arrays;
        uint32_t d_segments;
        float d_thickness, d_transparency, d_color[3];
        char info[88];
    } h;
    static_assert(sizeof(Header) == 128, "HAIR header is 128 bytes");
    f.read((char*)&h, sizeof(h));
    if (!f) throw IoError("cannot read hair file header");
    if (memcmp(h.sig, "HAIR", 4) != 0) throw std::invalid_argument("hair file has wrong signature");
    std::vector<uint16_t> segs;
    std::vector<float> pts, thick;
    if (h.arrays & 1) { segs.resize(h.hair_count); f.read((char*)segs.data(), 2 * (size_t)h.hair_count); if (!f) throw IoError("cannot read hair segments"); }
    if (h.arrays & 2) { pts.resize(3 * (size_t)h.point_count); f.read((char*)pts.data(), 12 * (size_t)h.point_count); if (!f) throw IoError("cannot read hair points"); }
    else throw std::invalid_argument("hair file has no points array");
    if (h.arrays & 4) { thick.resize(h.point_count); f.read((char*)thick.data(), 4 * (size_t)h.point_count); if (!f) throw IoError("cannot read hair thickness"); }

    // strands -> Catmull-Rom control points with mirrored phantom endpoints, radius 0.2 * thickness
    geo.cps.clear(); geo.seg_cp.clear(); geo.seg_strand.clear();
    geo.cps.reserve(h.point_count + 2 * (size_t)h.hair_count);
    for (int k = 0; k < 3; ++k) { geo.hair_min[k] = 0.f; geo.hair_max[k] = 0.f; }
    size_t p = 0;
    int strands = 0;
    for (uint32_t s = 0; s < h.hair_count && p < h.point_count; ++s) {
        const int nseg = segs.empty() ? (int)h.d_segments : (int)segs[s];
        const int npts = nseg + 1;
        if (p + npts > h.point_count) break;
        auto P = [&](size_t i, int k) { return pts[3 * i + k]; };
        auto W = [&](size_t i) { return 0.2f * (thick.empty() ? h.d_thickness : thick[i]); };
        for (int i = 0; i < npts; ++i)
            for (int k = 0; k < 3; ++k) {
                geo.hair_max[k] = std::max(geo.hair_max[k], P(p + i, k));
                geo.hair_min[k] = std::min(geo.hair_min[k], P(p + i, k));
            }
        if (nseg >= 1) {
            const int base = (int)geo.cps.size();
            geo.cps.push_back(F4{P(p, 0) + (P(p, 0) - P(p + 1, 0)), P(p, 1) + (P(p, 1) - P(p + 1, 1)), P(p, 2) + (P(p, 2) - P(p + 1, 2)), W(p)});
            for (int i = 0; i < npts; ++i) geo.cps.push_back(F4{P(p + i, 0), P(p + i, 1), P(p + i, 2), W(p + i)});
            const size_t l = p + npts - 1;
            geo.cps.push_back(F4{P(l, 0) + (P(l, 0) - P(l - 1, 0)), P(l, 1) + (P(l, 1) - P(l - 1, 1)), P(l, 2) + (P(l, 2) - P(l - 1, 2)), W(l)});
            for (int i = 0; i < nseg; ++i) { geo.seg_cp.push_back(base + i); geo.seg_strand.push_back(strands); }
        }
        p += npts;
        ++strands;
    }
    geo.num_strands = (int)h.hair_count;
}

// =====================================================================================
// .obj
// =====================================================================================
void load_obj_file(const std::string& path, HostGeometry& geo) {
    std::ifstream f(path);
    if (!f) throw IoError("Could not read OBJ model from " + path);
    std::vector<float> v, vn;
    std::string mtllib;
    geo.tri_verts.clear(); geo.tri_normals.clear(); geo.tri_uv.clear();
    std::string line;
    struct Idx { int v, t, n; };
    std::vector<Idx> face;
    std::vector<std::string> used_mtls;
    while (std::getline(f, line)) {
        const char* s = line.c_str();
        while (*s == ' ' || *s == '\t') ++s;
        if (s[0] == 'v' && s[1] == ' ') {
            float x, y, z;
            if (sscanf(s + 2, "%f %f %f", &x, &y, &z) == 3) { v.push_back(x); v.push_back(y); v.push_back(z); }
        } else if (s[0] == 'v' && s[1] == 'n') {
            float x, y, z;
            if (sscanf(s + 3, "%f %f %f", &x, &y, &z) == 3) { vn.push_back(x); vn.push_back(y); vn.push_back(z); }
        } else if (s[0] == 'f' && s[1] == ' ') {
            face.clear();
            const char* c = s + 2;
            while (*c) {
                while (*c == ' ' || *c == '\t' || *c == '\r') ++c;
                if (!*c) break;
                Idx id{0, 0, 0};
                const char* before = c;
                id.v = (int)strtol(c, (char**)&c, 10);
                if (c == before) throw std::invalid_argument("malformed face record in " + path + ": " + line);
                if (*c == '/') {
                    ++c;
                    if (*c != '/') id.t = (int)strtol(c, (char**)&c, 10);
                    if (*c == '/') { ++c; id.n = (int)strtol(c, (char**)&c, 10); }
                }
                face.push_back(id);
            }
            const int nv = (int)(v.size() / 3), nn = (int)(vn.size() / 3);
            auto fix = [](int i, int n) { return i > 0 ? i - 1 : (i < 0 ? n + i : -1); };
            for (size_t k = 1; k + 1 < face.size(); ++k) {   // fan triangulation
                const Idx tri[3] = {face[0], face[k], face[k + 1]};
                for (int c3 = 0; c3 < 3; ++c3) {
                    int vi = fix(tri[c3].v, nv), ni = fix(tri[c3].n, nn);
                    if (vi < 0 || vi >= nv) throw std::invalid_argument("invalid triangle indices");
                    geo.tri_verts.push_back(F4{v[3 * vi], v[3 * vi + 1], v[3 * vi + 2], 0.f});
                    if (ni >= 0 && ni < nn) geo.tri_normals.push_back(F4{vn[3 * ni], vn[3 * ni + 1], vn[3 * ni + 2], 0.f});
                    else geo.tri_normals.push_back(F4{0.f, 0.f, 0.f, 0.f});
                }
            }
        } else if (strncmp(s, "usemtl", 6) == 0) {
            std::string nm = s + 6;
            while (!nm.empty() && (nm.front() == ' ' || nm.front() == '\t')) nm.erase(nm.begin());
            while (!nm.empty() && (nm.back() == '\r' || nm.back() == ' ')) nm.pop_back();
            if (std::find(used_mtls.begin(), used_mtls.end(), nm) == used_mtls.end()) used_mtls.push_back(nm);
        } else if (strncmp(s, "mtllib", 6) == 0) {
            mtllib = s + 7;
            while (!mtllib.empty() && (mtllib.back() == '\r' || mtllib.back() == ' ')) mtllib.pop_back();
        }
    }
    // missing normals -> geometric normal
    for (size_t t = 0; t + 2 < geo.tri_verts.size(); t += 3) {
        F4& n0 = geo.tri_normals[t];
        if (n0.x == 0.f && n0.y == 0.f && n0.z == 0.f) {
            const F4 &a = geo.tri_verts[t], &b = geo.tri_verts[t + 1], &c = geo.tri_verts[t + 2];
            float e1[3] = {b.x - a.x, b.y - a.y, b.z - a.z}, e2[3] = {c.x - a.x, c.y - a.y, c.z - a.z};
            F4 n{e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0], 0.f};
            geo.tri_normals[t] = geo.tri_normals[t + 1] = geo.tri_normals[t + 2] = n;
        }
    }
    // material: the reference builds one mesh per material id, each with its own Kd (model.cpp:258-330), and
    // alpha fixed at 1 (model.cpp:322).  Here the surface carries ONE Kd: the materials the faces use must
    // agree on it (the shipped heads have a single material without Kd -> tinyobj's default 0).
    geo.kd[0] = geo.kd[1] = geo.kd[2] = 0.f;
    geo.surf_alpha = 1.f;
    if (!mtllib.empty()) {
        std::string dir = path.substr(0, path.rfind('/') + 1);
        std::ifstream m(dir + mtllib);
        struct Mtl { std::string name; float kd[3]; };
        std::vector<Mtl> mtls;
        while (m && std::getline(m, line)) {
            const char* s = line.c_str();
            while (*s == ' ' || *s == '\t') ++s;
            if (strncmp(s, "newmtl", 6) == 0) {
                std::string nm = s + 6;
                while (!nm.empty() && (nm.front() == ' ' || nm.front() == '\t')) nm.erase(nm.begin());
                while (!nm.empty() && (nm.back() == '\r' || nm.back() == ' ')) nm.pop_back();
                mtls.push_back(Mtl{nm, {0.f, 0.f, 0.f}});
            } else if (s[0] == 'K' && s[1] == 'd' && s[2] == ' ' && !mtls.empty()) {
                sscanf(s + 3, "%f %f %f", &mtls.back().kd[0], &mtls.back().kd[1], &mtls.back().kd[2]);
            }
        }
        const Mtl* chosen = nullptr;
        for (const std::string& u : used_mtls)
            for (const Mtl& mt : mtls) {
                if (mt.name != u) continue;
                if (chosen && (chosen->kd[0] != mt.kd[0] || chosen->kd[1] != mt.kd[1] || chosen->kd[2] != mt.kd[2]))
                    throw std::invalid_argument("OBJ uses several materials with different Kd (" + chosen->name + ", " + mt.name +
                                                "): per-material surface colours are not supported");
                if (!chosen) chosen = &mt;
            }
        if (!chosen && !mtls.empty()) chosen = &mtls[0];
        if (chosen) for (int k = 0; k < 3; ++k) geo.kd[k] = chosen->kd[k];
    }
}

// =====================================================================================
// OpenEXR reading (single-part scanline; HALF/FLOAT/UINT channels)
// =====================================================================================
void piz_decompress(const uint8_t* src, size_t src_len, uint16_t* out, size_t out_count,
                    const std::vector<int>& chan_u16_per_pixel, int nx, int ny);

namespace {
float half_to_float(uint16_t h) {
    uint32_t s = (h >> 15) & 1, e = (h >> 10) & 0x1f, m = h & 0x3ff;
    uint32_t bits;
    if (e == 0) {
        if (m == 0) bits = s << 31;
        else {
            int ee = -1;
            do { ee++; m <<= 1; } while (!(m & 0x400));
            bits = (s << 31) | ((uint32_t)(127 - 15 - ee) << 23) | ((m & 0x3ff) << 13);
        }
    } else if (e == 31) bits = (s << 31) | 0x7f800000u | (m << 13);
    else bits = (s << 31) | ((e + 112) << 23) | (m << 13);
    float f; memcpy(&f, &bits, 4);
    return f;
}

void zip_reconstruct(std::vector<uint8_t>& buf, std::vector<uint8_t>& tmp) {
    // predictor
    for (size_t i = 1; i < buf.size(); ++i) buf[i] = (uint8_t)(buf[i - 1] + buf[i] - 128);
    // de-interleave
    tmp.resize(buf.size());
    size_t half = (buf.size() + 1) / 2;
    const uint8_t* t1 = buf.data();
    const uint8_t* t2 = buf.data() + half;
    size_t o = 0;
    for (;;) {
        if (o < buf.size()) tmp[o++] = *t1++; else break;
        if (o < buf.size()) tmp[o++] = *t2++; else break;
    }
    buf.swap(tmp);
}
}  // namespace

void load_exr_rgba(const std::string& path, std::vector<float>& rgba, int& w, int& h) {
    std::string data = read_file(path);
    const uint8_t* d = (const uint8_t*)data.data();
    const size_t n = data.size();
    if (n < 8 || d[0] != 0x76 || d[1] != 0x2f || d[2] != 0x31 || d[3] != 0x01) throw std::invalid_argument("not an OpenEXR file: " + path);
    uint32_t version; memcpy(&version, d + 4, 4);
    if (version & 0x200) throw std::invalid_argument("tiled EXR files are not supported");
    if (version & 0x1800) throw std::invalid_argument("multi-part / deep EXR files are not supported");
    size_t p = 8;
    struct Chan { std::string name; int type; int xs, ys; };
    std::vector<Chan> chans;
    int compression = -1, line_order = 0;
    int dw[4] = {0, 0, -1, -1};
    auto rd_str = [&]() { std::string s; while (p < n && d[p]) s += (char)d[p++]; ++p; return s; };
    for (;;) {
        if (p >= n) throw std::invalid_argument("truncated EXR header");
        if (d[p] == 0) { ++p; break; }
        std::string name = rd_str(), type = rd_str();
        if (p + 4 > n) throw std::invalid_argument("truncated EXR header");
        uint32_t size; memcpy(&size, d + p, 4); p += 4;
        if (p + (size_t)size > n) throw std::invalid_argument("truncated EXR header");
        if (name == "channels") {
            size_t q = p;
            while (q < p + size && d[q]) {
                Chan c;
                while (q < p + size && d[q]) c.name += (char)d[q++];
                ++q;
                if (q + 16 > p + size) throw std::invalid_argument("truncated EXR channel list");
                int32_t t; memcpy(&t, d + q, 4); c.type = t; q += 8;   // type + pLinear/reserved
                int32_t xs, ys; memcpy(&xs, d + q, 4); memcpy(&ys, d + q + 4, 4); q += 8;
                c.xs = xs; c.ys = ys;
                if (xs != 1 || ys != 1) throw std::invalid_argument("sub-sampled EXR channels are not supported");
                chans.push_back(c);
            }
        } else if (name == "compression" && size >= 1) compression = d[p];
        else if (name == "dataWindow" && size >= 16) memcpy(dw, d + p, 16);
        else if (name == "lineOrder" && size >= 1) line_order = d[p];
        p += size;
    }
    (void)line_order;
    w = dw[2] - dw[0] + 1; h = dw[3] - dw[1] + 1;
    if (dw[2] < dw[0] || dw[3] < dw[1] || (long long)dw[2] - dw[0] >= 65536 || (long long)dw[3] - dw[1] >= 65536 || chans.empty())
        throw std::invalid_argument("bad EXR header");
    int lines_per_block;
    switch (compression) {
        case 0: case 1: case 2: lines_per_block = 1; break;   // NONE, RLE, ZIPS
        case 3: lines_per_block = 16; break;                  // ZIP
        case 4: lines_per_block = 32; break;                  // PIZ
        default: throw std::invalid_argument("EXR compression " + std::to_string(compression) + " is not supported (NONE/ZIPS/ZIP/PIZ only)");
    }
    if (compression == 1) throw std::invalid_argument("EXR RLE compression is not supported");
    const int nblocks = (h + lines_per_block - 1) / lines_per_block;
    if (p + 8 * (size_t)nblocks > n) throw std::invalid_argument("truncated EXR offset table");
    std::vector<uint64_t> offsets(nblocks);
    memcpy(offsets.data(), d + p, 8 * (size_t)nblocks);

    std::vector<int> bpp(chans.size()), u16pp(chans.size());
    size_t line_bytes = 0;
    for (size_t c = 0; c < chans.size(); ++c) {
        bpp[c] = chans[c].type == 1 ? 2 : 4;
        u16pp[c] = bpp[c] / 2;
        line_bytes += (size_t)bpp[c] * w;
    }
    // map channels to RGBA
    int slot_of[4] = {-1, -1, -1, -1};
    for (size_t c = 0; c < chans.size(); ++c) {
        const std::string& nm = chans[c].name;
        std::string base = nm.substr(nm.rfind('.') == std::string::npos ? 0 : nm.rfind('.') + 1);
        if (base == "R") slot_of[0] = (int)c; else if (base == "G") slot_of[1] = (int)c;
        else if (base == "B") slot_of[2] = (int)c; else if (base == "A") slot_of[3] = (int)c;
    }
    if (slot_of[0] < 0 && chans.size() == 1) slot_of[0] = slot_of[1] = slot_of[2] = 0;   // luminance-only
    rgba.assign((size_t)w * h * 4, 0.f);
    for (size_t i = 0; i < (size_t)w * h; ++i) rgba[4 * i + 3] = 1.f;

    std::vector<uint8_t> raw, tmp;
    for (int b = 0; b < nblocks; ++b) {
        size_t o = (size_t)offsets[b];
        if (o + 8 > n) throw std::invalid_argument("bad EXR block offset");
        int32_t y0, len; memcpy(&y0, d + o, 4); memcpy(&len, d + o + 4, 4);
        o += 8;
        if (len < 0 || o + (size_t)len > n) throw std::invalid_argument("bad EXR block size");
        const long long row0_ll = (long long)y0 - dw[1];
        if (row0_ll < 0 || row0_ll >= h) throw std::invalid_argument("EXR block lies outside the data window");
        const int row0 = (int)row0_ll;
        const int rows = std::min(lines_per_block, h - row0);
        const size_t expect = line_bytes * rows;
        raw.resize(expect);
        if ((size_t)len == expect) {
            memcpy(raw.data(), d + o, expect);
        } else if (compression == 2 || compression == 3) {
            uLongf dst = (uLongf)expect;
            if (uncompress(raw.data(), &dst, d + o, (uLong)len) != Z_OK || dst != expect) throw std::invalid_argument("EXR zlib block is corrupt");
            zip_reconstruct(raw, tmp);
        } else if (compression == 4) {
            piz_decompress(d + o, (size_t)len, (uint16_t*)raw.data(), expect / 2, u16pp, w, rows);
        } else {
            throw std::invalid_argument("EXR block size mismatch");
        }
        // block layout: for each scanline, for each channel (file order), w samples
        const uint8_t* src = raw.data();
        for (int r = 0; r < rows; ++r) {
            float* dst_row = rgba.data() + 4 * (size_t)(row0 + r) * w;
            for (size_t c = 0; c < chans.size(); ++c) {
                for (int slot = 0; slot < 4; ++slot) {
                    if (slot_of[slot] != (int)c) continue;
                    for (int x = 0; x < w; ++x) {
                        float v;
                        if (chans[c].type == 1) { uint16_t hv; memcpy(&hv, src + 2 * (size_t)x, 2); v = half_to_float(hv); }
                        else if (chans[c].type == 2) memcpy(&v, src + 4 * (size_t)x, 4);
                        else { uint32_t u; memcpy(&u, src + 4 * (size_t)x, 4); v = (float)u; }
                        dst_row[4 * (size_t)x + slot] = v;
                    }
                }
                src += (size_t)bpp[c] * w;
            }
        }
    }
}

// =====================================================================================
// Writers
// =====================================================================================
namespace {
uint32_t crc_table[256];
bool crc_ready = false;
uint32_t crc32_update(uint32_t c, const uint8_t* buf, size_t len) {
    if (!crc_ready) {
        for (uint32_t n = 0; n < 256; ++n) {
            uint32_t k = n;
            for (int i = 0; i < 8; ++i) k = (k & 1) ? 0xedb88320u ^ (k >> 1) : k >> 1;
            crc_table[n] = k;
        }
        crc_ready = true;
    }
    for (size_t i = 0; i < len; ++i) c = crc_table[(c ^ buf[i]) & 0xff] ^ (c >> 8);
    return c;
}
void put_be32(std::vector<uint8_t>& v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }
void png_chunk(std::vector<uint8_t>& out, const char* type, const std::vector<uint8_t>& payload) {
    put_be32(out, (uint32_t)payload.size());
    size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    out.insert(out.end(), payload.begin(), payload.end());
    uint32_t c = crc32_update(0xffffffffu, out.data() + start, out.size() - start) ^ 0xffffffffu;
    put_be32(out, c);
}
}  // namespace

void write_png_flipped(const std::string& path, const uint32_t* fb, int w, int h) {
    std::vector<uint8_t> raw;
    raw.reserve(((size_t)w * 4 + 1) * h);
    for (int y = 0; y < h; ++y) {
        const uint32_t* line = fb + (size_t)(h - 1 - y) * w;
        raw.push_back(0);   // filter: none
        for (int x = 0; x < w; ++x) {
            uint32_t px = line[x] | (0xffu << 24);
            raw.push_back(px & 0xff); raw.push_back((px >> 8) & 0xff); raw.push_back((px >> 16) & 0xff); raw.push_back(px >> 24);
        }
    }
    uLongf bound = compressBound((uLong)raw.size());
    std::vector<uint8_t> z(bound);
    if (compress2(z.data(), &bound, raw.data(), (uLong)raw.size(), 6) != Z_OK) throw std::runtime_error("png: deflate failed");
    z.resize(bound);
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::vector<uint8_t> ihdr;
    put_be32(ihdr, (uint32_t)w); put_be32(ihdr, (uint32_t)h);
    ihdr.push_back(8); ihdr.push_back(6); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    png_chunk(out, "IHDR", ihdr);
    png_chunk(out, "IDAT", z);
    png_chunk(out, "IEND", {});
    std::ofstream f(path, std::ios::binary);
    if (!f) throw IoError("cannot write " + path);
    f.write((const char*)out.data(), (std::streamsize)out.size());
}

void write_exr_flipped(const std::string& path, const float* rgba, int w, int h) {
    std::vector<uint8_t> out;
    auto put = [&](const void* p, size_t n) { out.insert(out.end(), (const uint8_t*)p, (const uint8_t*)p + n); };
    auto put_str = [&](const char* s) { put(s, strlen(s) + 1); };
    auto put_i32 = [&](int32_t v) { put(&v, 4); };
    auto put_f32 = [&](float v) { put(&v, 4); };
    const uint8_t magic[8] = {0x76, 0x2f, 0x31, 0x01, 2, 0, 0, 0};
    put(magic, 8);
    // channels (alphabetical): A B G R, all FLOAT
    put_str("channels"); put_str("chlist");
    put_i32(4 * 18 + 1);
    for (const char* c : {"A", "B", "G", "R"}) { put_str(c); put_i32(2); put_i32(0); put_i32(1); put_i32(1); }
    out.push_back(0);
    put_str("compression"); put_str("compression"); put_i32(1); out.push_back(0);
    put_str("dataWindow"); put_str("box2i"); put_i32(16); put_i32(0); put_i32(0); put_i32(w - 1); put_i32(h - 1);
    put_str("displayWindow"); put_str("box2i"); put_i32(16); put_i32(0); put_i32(0); put_i32(w - 1); put_i32(h - 1);
    put_str("lineOrder"); put_str("lineOrder"); put_i32(1); out.push_back(0);
    put_str("pixelAspectRatio"); put_str("float"); put_i32(4); put_f32(1.f);
    put_str("screenWindowCenter"); put_str("v2f"); put_i32(8); put_f32(0.f); put_f32(0.f);
    put_str("screenWindowWidth"); put_str("float"); put_i32(4); put_f32(1.f);
    out.push_back(0);
    const size_t table = out.size();
    const size_t line_bytes = (size_t)w * 16;
    out.resize(table + 8 * (size_t)h);
    for (int y = 0; y < h; ++y) {
        uint64_t off = out.size();
        memcpy(out.data() + table + 8 * (size_t)y, &off, 8);
        put_i32(y); put_i32((int32_t)line_bytes);
        const float* src = rgba + 4 * (size_t)(h - 1 - y) * w;   // vertical flip
        for (int c : {3, 2, 1, 0})                               // A B G R planes
            for (int x = 0; x < w; ++x) put_f32(src[4 * (size_t)x + c]);
    }
    std::ofstream f(path, std::ios::binary);
    if (!f) throw IoError("cannot write " + path);
    f.write((const char*)out.data(), (std::streamsize)out.size());
}

// =====================================================================================
// config.json
// =====================================================================================
namespace {
float jf(const Json& j) { return (float)j.number(); }
void vec3(const Json& j, float* o) {
    if (j.type != Json::Array || j.arr.size() < 3) throw std::invalid_argument("json: expected a 3-vector");
    for (int k = 0; k < 3; ++k) o[k] = jf(j.arr[k]);
}
}  // namespace

void load_scene_file(const std::string& config_path, HostScene& s) {
    Json cfg = parse_json_file(config_path);
    size_t slash = config_path.rfind('/');
    s.base_dir = slash == std::string::npos ? "." : config_path.substr(0, slash);

    bool is_obj = false, is_hair = false;
    if (const Json* surf = cfg.find("surface"))
        if (const Json* g = surf->find("geometry")) {
            load_obj_file(resolve_scene_path(g->string(), s.base_dir), s.geo);
            is_obj = true;
        }
    const Json* hair = cfg.find("hair");
    if (hair)
        if (const Json* g = hair->find("geometry")) {
            load_hair_file(resolve_scene_path(g->string(), s.base_dir), s.geo);
            is_hair = true;
        }
    if (!is_obj && !is_hair) throw std::invalid_argument("Either hair or surface must be defined!");

    const Json* cam = cfg.find("camera");
    if (!cam) throw std::invalid_argument("Camera must be defined!");
    vec3(cam->at("from"), s.cam_from); vec3(cam->at("to"), s.cam_to); vec3(cam->at("up"), s.cam_up);
    s.cos_fovy = jf(cam->at("cos_fovy"));

    if (hair && hair->has("type")) {
        if (const Json* v = hair->find("sigma_a")) vec3(*v, s.sigma_a);
        if (const Json* v = hair->find("beta_m")) s.beta_m = jf(*v);
        if (const Json* v = hair->find("beta_n")) s.beta_n = jf(*v);
        if (const Json* v = hair->find("alpha")) s.alpha = 3.14159f * jf(*v) / 180.f;
        if (const Json* v = hair->find("Gain R")) s.gains[0] = jf(*v);
        if (const Json* v = hair->find("Gain TT")) s.gains[1] = jf(*v);
        if (const Json* v = hair->find("Gain TRT")) s.gains[2] = jf(*v);
        if (const Json* v = hair->find("Gain TRRT")) s.gains[3] = jf(*v);
    }

    const Json* lights = cfg.find("lights");
    s.has_env = false;
    if (lights)
        if (const Json* env = lights->find("environment"))
            if (const Json* exr = env->find("exr")) {
                load_exr_rgba(resolve_scene_path(exr->string(), s.base_dir), s.env, s.env_w, s.env_h);
                s.has_env = true;
                s.env_scale = jf(env->at("scale"));
                s.env_rot = env->has("rotation") ? jf(env->at("rotation")) : 0.f;
            }
    bool is_dir = false;
    if (lights)
        if (const Json* dl = lights->find("directional")) {
            for (const Json& l : dl->arr) {
                float from[3], emit[3];
                vec3(l.at("from"), from); vec3(l.at("emit"), emit);
                float r = 1.f / sqrtf(from[0] * from[0] + from[1] * from[1] + from[2] * from[2]);
                for (int k = 0; k < 3; ++k) s.dl_from.push_back(from[k] * r);
                for (int k = 0; k < 3; ++k) s.dl_emit.push_back(emit[k]);
            }
            is_dir = true;
        }
    if (!is_dir && !s.has_env) throw std::invalid_argument("Either directional or environment light must be defined!");

    const Json* integ = cfg.find("integrator");
    if (!integ) throw std::invalid_argument("Integrator must be defined!");
    s.spp = (int)integ->at("spp").number();
    s.path_v1 = (int)integ->at("path_v1").number();
    s.path_v2 = (int)integ->at("path_v2").number();
    s.width = (int)integ->at("width").number();
    s.height = (int)integ->at("height").number();
    s.image_output = integ->at("image_output").string();
    s.stats_output = integ->at("stats_output").string();
    s.mis = integ->at("MIS").boolean();
    s.env_pdf = integ->at("ENV_PDF").boolean();
    // scene.cpp:302-306 rejects W*H > 2048^2 and W*H % 128 != 0.  The 2K*2K limit is applied PER GPU BAND here
    // (hm_renderer_create: each band's pixel count), so that a 4096^2 frame renders on >= 4 row bands
    // (BASELINE config 5); the frame itself is bounded by 8192^2.
    if (s.width <= 0 || s.height <= 0 || (int64_t)s.width * s.height > (int64_t)8192 * 8192 || ((int64_t)s.width * s.height) % 128 != 0)
        throw std::invalid_argument("Image size must be a positive multiple of 128 pixels, at most 8K*8K (2K*2K per GPU band)!");

    if (const Json* t = cfg.find("tcnn")) {
        if (const Json* c = t->find("config")) {
            try { s.tcnn_config = resolve_scene_path(c->string(), s.base_dir); }
            catch (const IoError&) { s.tcnn_config.clear(); }
        }
        s.tcnn_train = t->has("init_train") ? t->at("init_train").boolean() : false;
        if (const Json* wts = t->find("init_weights")) s.tcnn_weights = wts->string();
    }
}

}  // namespace hm
