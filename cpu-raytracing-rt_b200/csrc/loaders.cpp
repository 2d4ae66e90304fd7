// loaders.cpp — host: the two scene formats of the reference and its PPM writer.
//   text : reference src/scene_parser.rs:5-85, defaults of src/scene.rs:92-106, :167-191
//   glTF : reference src/gltf/parser.rs (field names/defaults), src/gltf/scene_builder.rs:9-388
//   PPM  : reference src/ppm.rs:4-19
// Where the reference panics (unwrap / expect / assert!) these throw std::runtime_error, which the
// C ABI turns into RT_ERR_PARSE / RT_ERR_IO.
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <istream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>

#include "scene_host.hpp"

namespace rt {

void RawSceneStorage::link() {
    raw.n_materials = (uint32_t)materials.size();
    raw.n_shapes = (uint32_t)shapes.size();
    raw.materials = materials.data();
    raw.shapes = shapes.data();
    raw.n_triangles = tri_pos.size() / 9;
    raw.tri_pos = tri_pos.empty() ? nullptr : tri_pos.data();
    raw.tri_nrm = tri_nrm.empty() ? nullptr : tri_nrm.data();
    raw.tri_mat = tri_mat.empty() ? nullptr : tri_mat.data();
    raw.tri_xform = tri_xform.empty() ? nullptr : tri_xform.data();
}

namespace {

[[noreturn]] void fail(const std::string &m) { throw std::runtime_error(m); }

double to_f64(const std::string &tok) {         // str::parse::<f64>
    if (tok.empty() || tok.find_first_of("xX") != std::string::npos) fail("bad number '" + tok + "'");
    char *end = nullptr;
    const double v = std::strtod(tok.c_str(), &end);
    if (end != tok.c_str() + tok.size()) fail("bad number '" + tok + "'");
    return v;
}
uint64_t to_uint(const std::string &tok, uint64_t max) {    // usize / u8 parse
    size_t i = (tok.size() > 1 && tok[0] == '+') ? 1 : 0;
    if (i >= tok.size()) fail("bad integer '" + tok + "'");
    uint64_t v = 0;
    for (; i < tok.size(); ++i) {
        if (tok[i] < '0' || tok[i] > '9') fail("bad integer '" + tok + "'");
        v = v * 10 + (uint64_t)(tok[i] - '0');
        if (v > max) fail("integer out of range '" + tok + "'");
    }
    return v;
}
d3 unit_or(const double *v) { return unit3(mk3(v[0], v[1], v[2])); }

struct TextPrim {
    int type = -1;                  // 0 plane 1 box 2 ellipsoid 3 triangle
    double param[9] = {0};
    bool has_material = false, has_ior = false, has_pos = false, has_rot = false;
    int material = RT_MAT_DIFFUSE;
    double ior = 1.0, pos[3] = {0, 0, 0}, rot[4] = {0, 0, 0, 1}, color[3] = {0, 0, 0}, emission[3] = {0, 0, 0};
};

}  // namespace

void parse_text_scene(std::istream &in, RawSceneStorage &out) {
    std::vector<TextPrim> prims;
    bool has_dims = false;
    rt_raw_scene &raw = out.raw;
    raw = rt_raw_scene{};
    raw.samples = 64;               // src/scene.rs:187
    raw.ray_depth = 16;             // src/scene.rs:185
    raw.fov_kind = RT_FOV_X;
    raw.fov = kPi / 2.0;            // src/scene.rs:174
    raw.tri_mode = RT_TRI_MODEL;
    double right[3] = {1, 0, 0}, up[3] = {0, 1, 0}, forward[3] = {0, 0, 1};
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ls(line);
        std::vector<std::string> tok;
        for (std::string t; ls >> t;) tok.push_back(t);
        if (tok.empty()) continue;
        size_t pos = 1;
        auto next = [&]() -> const std::string & {
            if (pos >= tok.size()) fail("missing value after " + tok[0]);
            return tok[pos++];
        };
        auto vec = [&](double *dst, int n) { for (int i = 0; i < n; ++i) dst[i] = to_f64(next()); };
        auto last = [&]() -> TextPrim & {
            if (prims.empty()) fail(tok[0] + " before any NEW_PRIMITIVE");
            return prims.back();
        };
        const std::string &k = tok[0];
        if (k == "NEW_PRIMITIVE") prims.emplace_back();
        else if (k == "BOX") { TextPrim &p = last(); p.type = 1; vec(p.param, 3); }
        else if (k == "PLANE") { TextPrim &p = last(); p.type = 0; vec(p.param, 3); }
        else if (k == "ELLIPSOID") { TextPrim &p = last(); p.type = 2; vec(p.param, 3); }
        else if (k == "TRIANGLE") { TextPrim &p = last(); p.type = 3; vec(p.param, 9); }
        else if (k == "POSITION") { TextPrim &p = last(); vec(p.pos, 3); p.has_pos = true; }
        else if (k == "ROTATION") { TextPrim &p = last(); vec(p.rot, 4); p.has_rot = true; }
        else if (k == "COLOR") vec(last().color, 3);
        else if (k == "EMISSION") vec(last().emission, 3);
        else if (k == "METALLIC") { last().material = RT_MAT_METALLIC; }
        else if (k == "DIELECTRIC") { last().material = RT_MAT_DIELECTRIC; }
        else if (k == "IOR") { TextPrim &p = last(); p.ior = to_f64(next()); p.has_ior = true; }
        else if (k == "CAMERA_POSITION") vec(raw.cam_position, 3);
        else if (k == "CAMERA_RIGHT") vec(right, 3);
        else if (k == "CAMERA_UP") vec(up, 3);
        else if (k == "CAMERA_FORWARD") vec(forward, 3);
        else if (k == "CAMERA_FOV_X") raw.fov = to_f64(next());
        else if (k == "DIMENSIONS") {
            raw.width = (uint32_t)to_uint(next(), 0xFFFFFFFFull);
            raw.height = (uint32_t)to_uint(next(), 0xFFFFFFFFull);
            has_dims = true;
        } else if (k == "RAY_DEPTH") raw.ray_depth = (uint32_t)to_uint(next(), 255);
        else if (k == "BG_COLOR") vec(raw.bg_color, 3);
        else if (k == "SAMPLES") raw.samples = (uint32_t)to_uint(next(), 0xFFFFFFFFull);
        // unknown keyword: skipped (src/scene_parser.rs:37)
    }
    if (!has_dims) fail("DIMENSIONS missing");               // dimensions.unwrap(), src/scene.rs:188
    const d3 r = unit_or(right), u = unit_or(up), f = unit_or(forward);     // src/scene.rs:171-173
    raw.cam_right[0] = r.x; raw.cam_right[1] = r.y; raw.cam_right[2] = r.z;
    raw.cam_up[0] = u.x; raw.cam_up[1] = u.y; raw.cam_up[2] = u.z;
    raw.cam_forward[0] = f.x; raw.cam_forward[1] = f.y; raw.cam_forward[2] = f.z;
    for (const TextPrim &p : prims) {
        if (p.type < 0) fail("primitive without a shape");   // prim_type.unwrap(), src/scene.rs:201
        if (p.material == RT_MAT_DIELECTRIC && !p.has_ior) fail("DIELECTRIC without IOR");  // src/scene.rs:97
        rt_material m{};
        m.kind = p.material;
        m.ior = p.ior;
        std::memcpy(m.color, p.color, sizeof m.color);
        std::memcpy(m.emission, p.emission, sizeof m.emission);
        out.materials.push_back(m);
        const uint32_t mi = (uint32_t)out.materials.size() - 1;
        if (p.type == 3) {
            out.tri_pos.insert(out.tri_pos.end(), p.param, p.param + 9);
            out.tri_mat.push_back(mi);
            out.tri_xform.insert(out.tri_xform.end(), p.pos, p.pos + 3);
            out.tri_xform.insert(out.tri_xform.end(), p.rot, p.rot + 4);
        } else {
            rt_shape s{};
            s.type = p.type == 0 ? RT_SHAPE_PLANE : (p.type == 1 ? RT_SHAPE_BOX : RT_SHAPE_ELLIPSOID);
            s.material = mi;
            std::memcpy(s.param, p.param, sizeof s.param);
            std::memcpy(s.position, p.pos, sizeof s.position);
            std::memcpy(s.rotation, p.rot, sizeof s.rotation);
            out.shapes.push_back(s);
        }
    }
    out.link();
}

// ------------------------------------------------------------------------------------------------
// Minimal JSON reader (the reference uses serde_json; unknown fields are ignored there too).
namespace {

struct Json {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;
    const Json *get(const char *key) const {
        if (kind != Obj) return nullptr;
        for (const auto &kv : obj)
            if (kv.first == key) return kv.second.kind == Null ? nullptr : &kv.second;
        return nullptr;
    }
};

struct JsonParser {
    const char *p, *end;
    void ws() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p; }
    [[noreturn]] void err(const char *m) { fail(std::string("can't parse glTF: ") + m); }
    Json value() {
        ws();
        if (p >= end) err("unexpected end");
        Json j;
        if (*p == '{') {
            j.kind = Json::Obj;
            ++p; ws();
            if (p < end && *p == '}') { ++p; return j; }
            for (;;) {
                ws();
                if (p >= end || *p != '"') err("expected key");
                std::string k = string();
                ws();
                if (p >= end || *p != ':') err("expected ':'");
                ++p;
                j.obj.emplace_back(std::move(k), value());
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == '}') { ++p; break; }
                err("expected ',' or '}'");
            }
        } else if (*p == '[') {
            j.kind = Json::Arr;
            ++p; ws();
            if (p < end && *p == ']') { ++p; return j; }
            for (;;) {
                j.arr.push_back(value());
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == ']') { ++p; break; }
                err("expected ',' or ']'");
            }
        } else if (*p == '"') {
            j.kind = Json::Str;
            j.str = string();
        } else if (!std::strncmp(p, "true", 4)) { j.kind = Json::Bool; j.b = true; p += 4; }
        else if (!std::strncmp(p, "false", 5)) { j.kind = Json::Bool; j.b = false; p += 5; }
        else if (!std::strncmp(p, "null", 4)) { j.kind = Json::Null; p += 4; }
        else {
            char *e = nullptr;
            j.kind = Json::Num;
            j.num = std::strtod(p, &e);
            if (e == p) err("bad value");
            p = e;
        }
        return j;
    }
    std::string string() {
        std::string s;
        ++p;
        while (p < end && *p != '"') {
            if (*p == '\\') {
                ++p;
                if (p >= end) err("bad escape");
                switch (*p) {
                    case 'n': s += '\n'; break;
                    case 't': s += '\t'; break;
                    case 'r': s += '\r'; break;
                    case 'b': s += '\b'; break;
                    case 'f': s += '\f'; break;
                    case 'u': {
                        if (end - p < 5) err("bad \\u");
                        unsigned cp = (unsigned)std::strtoul(std::string(p + 1, p + 5).c_str(), nullptr, 16);
                        p += 4;
                        if (cp < 0x80) s += (char)cp;
                        else if (cp < 0x800) { s += (char)(0xC0 | (cp >> 6)); s += (char)(0x80 | (cp & 0x3F)); }
                        else { s += (char)(0xE0 | (cp >> 12)); s += (char)(0x80 | ((cp >> 6) & 0x3F)); s += (char)(0x80 | (cp & 0x3F)); }
                        break;
                    }
                    default: s += *p;
                }
                ++p;
            } else {
                s += *p++;
            }
        }
        if (p >= end) err("unterminated string");
        ++p;
        return s;
    }
};

struct M4 {                         // column-major like cgmath Matrix4
    double c[4][4];
};
M4 m4_identity() {
    M4 m{};
    for (int i = 0; i < 4; ++i) m.c[i][i] = 1.0;
    return m;
}
M4 m4_mul(const M4 &a, const M4 &b) {           // Matrix4 * Matrix4
    M4 o;
    for (int j = 0; j < 4; ++j)
        for (int r = 0; r < 4; ++r)
            o.c[j][r] = ((a.c[0][r] * b.c[j][0] + a.c[1][r] * b.c[j][1]) + a.c[2][r] * b.c[j][2]) + a.c[3][r] * b.c[j][3];
    return o;
}
M4 m4_from_quat(double w, double x, double y, double z) {       // From<Quaternion> for Matrix4
    const double x2 = x + x, y2 = y + y, z2 = z + z;
    const double xx2 = x2 * x, xy2 = x2 * y, xz2 = x2 * z, yy2 = y2 * y, yz2 = y2 * z, zz2 = z2 * z;
    const double sy2 = y2 * w, sz2 = z2 * w, sx2 = x2 * w;
    M4 m{};
    m.c[0][0] = 1.0 - yy2 - zz2; m.c[0][1] = xy2 + sz2; m.c[0][2] = xz2 - sy2;
    m.c[1][0] = xy2 - sz2; m.c[1][1] = 1.0 - xx2 - zz2; m.c[1][2] = yz2 + sx2;
    m.c[2][0] = xz2 + sy2; m.c[2][1] = yz2 - sx2; m.c[2][2] = 1.0 - xx2 - yy2;
    m.c[3][3] = 1.0;
    return m;
}

std::vector<double> num_array(const Json *j, size_t expect, const char *what) {
    std::vector<double> v;
    if (!j) return v;
    if (j->kind != Json::Arr) fail(std::string(what) + ": expected array");
    for (const Json &e : j->arr) {
        if (e.kind != Json::Num) fail(std::string(what) + ": expected numbers");
        v.push_back(e.num);
    }
    if (v.size() != expect) fail(std::string(what) + ": wrong length");
    return v;
}
size_t index_of(const Json *j, const char *what) {
    if (!j || j->kind != Json::Num || j->num < 0 || j->num != (double)(size_t)j->num) fail(std::string(what) + ": expected index");
    return (size_t)j->num;
}
const Json &at(const Json *arr, size_t i, const char *what) {
    if (!arr || arr->kind != Json::Arr || i >= arr->arr.size()) fail(std::string(what) + ": index out of range");
    return arr->arr[i];
}

struct GNode {
    M4 trs;
    std::vector<size_t> children;
    bool has_camera = false, has_mesh = false;
    size_t mesh = 0;
};

std::string read_file(const std::string &path, const char *what) {
    std::ifstream f(path, std::ios::binary);
    if (!f) fail(std::string("Couldn't find or load ") + what + " '" + path + "'");
    std::ostringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

constexpr double kUnsignedShort = 5123, kUnsignedInt = 5125, kFloat = 5126;

struct Gltf {
    Json doc;
    std::vector<GNode> nodes;
    std::map<std::string, std::string> buffers;
    RawSceneStorage *out = nullptr;

    struct View { const std::string *bytes; size_t offset; size_t stride; bool has_stride; size_t count; };
    bool view_of(const Json &acc, View &v, const char *what) {
        const Json *bv = acc.get("bufferView");
        if (!bv) return false;
        const Json &view = at(doc.get("bufferViews"), index_of(bv, what), what);
        const Json &buf = at(doc.get("buffers"), index_of(view.get("buffer"), what), what);
        const Json *uri = buf.get("uri");
        if (!uri || uri->kind != Json::Str) fail("expected uri for buffer");
        v.bytes = &buffers.at(uri->str);
        const Json *bo = view.get("byteOffset"), *ao = acc.get("byteOffset"), *st = view.get("byteStride");
        v.offset = (bo ? index_of(bo, what) : 0) + (ao ? index_of(ao, what) : 0);
        v.has_stride = st != nullptr;
        v.stride = st ? index_of(st, what) : 0;
        v.count = index_of(acc.get("count"), what);
        return true;
    }
    static float f32_at(const std::string &b, size_t pos) {     // read_float, :358-361
        if (pos + 4 > b.size()) fail("buffer too short");
        float f;
        std::memcpy(&f, b.data() + pos, 4);
        return f;
    }
    void read_vec3(const Json &acc, std::vector<d3> &res, const char *what) {
        View v;
        if (!view_of(acc, v, what)) return;
        const Json *ct = acc.get("componentType"), *ty = acc.get("type");
        if (!ct || ct->num != kFloat) fail(std::string(what) + ": componentType must be FLOAT");
        if (!ty || ty->str != "VEC3") fail(std::string(what) + ": must have VEC3 type");
        const size_t stride = v.has_stride ? v.stride : 12;
        res.resize(v.count);
        size_t off = v.offset;
        for (size_t i = 0; i < v.count; ++i, off += stride)
            res[i] = mk3((double)f32_at(*v.bytes, off), (double)f32_at(*v.bytes, off + 4), (double)f32_at(*v.bytes, off + 8));
    }
    void read_indices(const Json &acc, std::vector<uint32_t> &res) {    // :237-267
        View v;
        if (!view_of(acc, v, "indices")) return;
        const Json *ct = acc.get("componentType"), *ty = acc.get("type");
        if (!ct || (ct->num != kUnsignedShort && ct->num != kUnsignedInt)) fail("indices: componentType must be UNSIGNED_SHORT or UNSIGNED_INT");
        if (!ty || ty->str != "SCALAR") fail("indices must have SCALAR type");
        const bool is16 = ct->num == kUnsignedShort;
        const size_t esz = is16 ? 2 : 4, stride = v.has_stride ? v.stride : esz;
        res.resize(v.count);
        size_t off = v.offset;
        for (size_t i = 0; i < v.count; ++i, off += stride) {
            if (off + esz > v.bytes->size()) fail("buffer too short");
            const unsigned char *b = (const unsigned char *)v.bytes->data() + off;
            res[i] = is16 ? (uint32_t)(b[0] | (b[1] << 8)) : ((uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24));
        }
    }
    rt_material metadata(const Json *mat) {                     // make_metadata, :227-235
        double color[4] = {1, 1, 1, 1}, emission[3] = {0, 0, 0}, strength = 1.0, metallic = 1.0;
        if (mat) {
            if (const Json *pbr = mat->get("pbrMetallicRoughness")) {
                auto c = num_array(pbr->get("baseColorFactor"), 4, "baseColorFactor");
                if (!c.empty()) std::copy(c.begin(), c.end(), color);
                if (const Json *mf = pbr->get("metallicFactor")) metallic = mf->num;
            }
            auto e = num_array(mat->get("emissiveFactor"), 3, "emissiveFactor");
            if (!e.empty()) std::copy(e.begin(), e.end(), emission);
            if (const Json *ext = mat->get("extensions"))
                if (const Json *es = ext->get("KHR_materials_emissive_strength")) {
                    const Json *s = es->get("emissive_strength");
                    if (!s || s->kind != Json::Num) fail("emissive_strength missing");
                    strength = s->num;
                }
        }
        rt_material m{};
        m.kind = color[3] < 1.0 ? RT_MAT_DIELECTRIC : (metallic > 0.0 ? RT_MAT_METALLIC : RT_MAT_DIFFUSE);
        m.ior = m.kind == RT_MAT_DIELECTRIC ? 1.5 : 1.0;
        for (int i = 0; i < 3; ++i) { m.color[i] = color[i]; m.emission[i] = emission[i] * strength; }
        return m;
    }
    void convert_primitive(const Json &prim, const M4 &trs) {   // :209-225
        const Json *mode = prim.get("mode");
        if (mode && mode->num != 4) fail("supported only triangles for primitive.mode");
        const Json *attrs = prim.get("attributes");
        if (!attrs) fail("primitive without attributes");
        if (!attrs->get("NORMAL")) fail("empty normals");
        std::vector<d3> verts, normals;
        read_vec3(at(doc.get("accessors"), index_of(attrs->get("POSITION"), "POSITION"), "POSITION"), verts, "vertices");
        read_vec3(at(doc.get("accessors"), index_of(attrs->get("NORMAL"), "NORMAL"), "NORMAL"), normals, "normals");
        const d3 c0 = mk3(trs.c[0][0], trs.c[0][1], trs.c[0][2]), c1 = mk3(trs.c[1][0], trs.c[1][1], trs.c[1][2]);
        const d3 c2 = mk3(trs.c[2][0], trs.c[2][1], trs.c[2][2]), c3 = mk3(trs.c[3][0], trs.c[3][1], trs.c[3][2]);
        for (d3 &v : verts) {                                   // trs * vec4(v, 1), :285-292
            const double w = ((trs.c[0][3] * v.x + trs.c[1][3] * v.y) + trs.c[2][3] * v.z) + trs.c[3][3] * 1.0;
            if (w != 1.0) fail("assertion failed: pos.w == 1.0");
            v = add(add(add(scale(c0, v.x), scale(c1, v.y)), scale(c2, v.z)), scale(c3, 1.0));
        }
        // cof(mat3(trs)), :363-388
        const double *col[3] = {trs.c[0], trs.c[1], trs.c[2]};
        d3 rc[3];
        for (int cc = 0; cc < 3; ++cc) {
            double r[3];
            for (int rr = 0; rr < 3; ++rr) {
                const int lc = cc == 0 ? 1 : 0, rcx = cc == 2 ? 1 : 2, tr = rr == 0 ? 1 : 0, br = rr == 2 ? 1 : 2;
                const double det = col[lc][tr] * col[rcx][br] - col[rcx][tr] * col[lc][br];
                r[rr] = ((cc + rr) & 1) ? -det : det;
            }
            rc[cc] = mk3(r[0], r[1], r[2]);
        }
        for (d3 &n : normals) n = unit3(add(add(scale(rc[0], n.x), scale(rc[1], n.y)), scale(rc[2], n.z)));
        if (verts.size() != normals.size()) fail("vertices.len() != normals.len()");
        std::vector<uint32_t> idx;
        if (const Json *ia = prim.get("indices")) {
            read_indices(at(doc.get("accessors"), index_of(ia, "indices"), "indices"), idx);
            if (idx.size() % 3) fail("indices.len() % 3 != 0");
        } else {
            if (verts.size() % 3) fail("vertices.len() % 3 != 0");
            idx.resize(verts.size());
            for (size_t i = 0; i < idx.size(); ++i) idx[i] = (uint32_t)i;
        }
        const Json *mi = prim.get("material");
        out->materials.push_back(metadata(mi ? &at(doc.get("materials"), index_of(mi, "material"), "material") : nullptr));
        const uint32_t m = (uint32_t)out->materials.size() - 1;
        const size_t base = out->tri_pos.size();
        out->tri_pos.resize(base + idx.size() * 3);
        out->tri_nrm.resize(base + idx.size() * 3);
        for (size_t i = 0; i < idx.size(); ++i) {
            if (idx[i] >= verts.size()) fail("vertex index out of range");
            const d3 &v = verts[idx[i]], &n = normals[idx[i]];
            double *pp = &out->tri_pos[base + i * 3], *pn = &out->tri_nrm[base + i * 3];
            pp[0] = v.x; pp[1] = v.y; pp[2] = v.z;
            pn[0] = n.x; pn[1] = n.y; pn[2] = n.z;
        }
        out->tri_mat.insert(out->tri_mat.end(), idx.size() / 3, m);
    }
    void convert_node(size_t i) {                               // :190-200
        if (i >= nodes.size()) fail("node index out of range");
        const GNode &n = nodes[i];
        if (n.has_mesh) {
            const Json &mesh = at(doc.get("meshes"), n.mesh, "mesh");
            const Json *prims = mesh.get("primitives");
            if (!prims || prims->kind != Json::Arr) fail("mesh without primitives");
            for (const Json &p : prims->arr) convert_primitive(p, n.trs);
        }
        for (size_t c : n.children) convert_node(c);
    }
    void propagate(size_t i, const M4 &parent) {                // :163-169
        if (i >= nodes.size()) fail("node index out of range");
        nodes[i].trs = m4_mul(parent, nodes[i].trs);
        const M4 mine = nodes[i].trs;
        for (size_t c : nodes[i].children) propagate(c, mine);
    }
};

}  // namespace

void load_gltf_scene(const std::string &path, uint32_t w, uint32_t h, uint32_t spp, RawSceneStorage &out) {
    const size_t slash = path.rfind('/');                       // bin_file_prefix, src/main.rs:77-83
    const std::string prefix = slash == std::string::npos ? "" : path.substr(0, slash + 1);
    const std::string text = read_file(path, "gltf file");
    Gltf g;
    g.out = &out;
    JsonParser jp{text.data(), text.data() + text.size()};
    g.doc = jp.value();
    if (g.doc.kind != Json::Obj) fail("can't parse glTF: not an object");
    if (const Json *ns = g.doc.get("nodes")) {
        if (ns->kind != Json::Arr) fail("nodes: expected array");
        for (const Json &n : ns->arr) {                         // Node::new + extract_trs, :97-123
            GNode gn;
            if (const Json *m = n.get("matrix")) {
                auto v = num_array(m, 16, "matrix");
                for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) gn.trs.c[c][r] = v[c * 4 + r];
            } else {
                auto t = num_array(n.get("translation"), 3, "translation");
                auto r = num_array(n.get("rotation"), 4, "rotation");
                auto s = num_array(n.get("scale"), 3, "scale");
                M4 mt = m4_identity();
                if (!t.empty()) { mt.c[3][0] = t[0]; mt.c[3][1] = t[1]; mt.c[3][2] = t[2]; }
                M4 mr = r.empty() ? m4_from_quat(1, 0, 0, 0) : m4_from_quat(r[3], r[0], r[1], r[2]);
                M4 ms = m4_identity();
                if (!s.empty()) { ms.c[0][0] = s[0]; ms.c[1][1] = s[1]; ms.c[2][2] = s[2]; }
                gn.trs = m4_mul(m4_mul(mt, mr), ms);
            }
            if (const Json *ch = n.get("children")) {
                if (ch->kind != Json::Arr) fail("children: expected array");
                for (const Json &c : ch->arr) gn.children.push_back(index_of(&c, "children"));
            }
            gn.has_camera = n.get("camera") != nullptr;
            if (const Json *m = n.get("mesh")) { gn.has_mesh = true; gn.mesh = index_of(m, "mesh"); }
            g.nodes.push_back(std::move(gn));
        }
    }
    const Json *scenes = g.doc.get("scenes");
    if (scenes)                                                 // propagate_trs: EVERY scene's roots, :155-161
        for (const Json &sc : scenes->arr) {
            const Json *roots = sc.get("nodes");
            if (!roots || roots->kind != Json::Arr) fail("scene without nodes");
            for (const Json &r : roots->arr) g.propagate(index_of(&r, "scene.nodes"), m4_identity());
        }
    if (const Json *bufs = g.doc.get("buffers"))                // load_buffers, :80-87
        for (const Json &b : bufs->arr) {
            const Json *uri = b.get("uri");
            if (!uri || uri->kind != Json::Str) fail("expected uri for buffer");
            g.buffers[uri->str] = read_file(prefix + uri->str, "buffer");
        }
    const Json *scene_idx = g.doc.get("scene");
    const Json &main_scene = at(scenes, scene_idx ? index_of(scene_idx, "scene") : 0, "scenes[scene]");
    for (const Json &r : main_scene.get("nodes")->arr) g.convert_node(index_of(&r, "scene.nodes"));   // :183-185

    const Json *cams = g.doc.get("cameras");                    // extract_camera_params, :57-78
    const Json *persp = nullptr;
    if (cams && cams->kind == Json::Arr && cams->arr.size() == 1) {
        const Json *ty = cams->arr[0].get("type");
        if (ty && ty->kind == Json::Str && ty->str == "perspective") persp = cams->arr[0].get("perspective");
    }
    if (!persp || !persp->get("yfov")) fail("Supported only single perspective camera");
    const GNode *cam_node = nullptr;
    for (const GNode &n : g.nodes)
        if (n.has_camera) {
            if (cam_node) fail("You must specify only one a node with the camera");
            cam_node = &n;
        }
    if (!cam_node) fail("You must specify a node with the camera");
    rt_raw_scene &raw = out.raw;
    raw = rt_raw_scene{};
    raw.width = w; raw.height = h; raw.samples = spp;
    raw.ray_depth = 8;                                          // :16
    raw.fov_kind = RT_FOV_Y;
    raw.fov = persp->get("yfov")->num;
    raw.tri_mode = RT_TRI_BAKED;
    for (int k = 0; k < 3; ++k) {
        raw.cam_position[k] = cam_node->trs.c[3][k];
        raw.cam_right[k] = cam_node->trs.c[0][k];
        raw.cam_up[k] = cam_node->trs.c[1][k];
        raw.cam_forward[k] = -cam_node->trs.c[2][k];
    }
    out.link();
}

// ------------------------------------------------------------------------------------------------
void save_ppm(const std::string &path, uint32_t w, uint32_t h, const double *rgb) {     // src/ppm.rs:4-19
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) fail("Cannot create output file '" + path + "'");
    std::fprintf(f, "P6\n%u %u\n255\n", w, h);
    std::vector<unsigned char> row((size_t)w * 3);
    for (uint32_t y = 0; y < h; ++y) {
        for (size_t i = 0; i < row.size(); ++i) {
            double v = rgb[(size_t)y * w * 3 + i];
            v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);            // f64::clamp(0,1); NaN -> `as u8` gives 0
            const double r = std::round(v * 255.0);
            row[i] = r == r ? (unsigned char)r : 0;
        }
        if (std::fwrite(row.data(), 1, row.size(), f) != row.size()) { std::fclose(f); fail("write failed"); }
    }
    std::fclose(f);
}

}  // namespace rt
