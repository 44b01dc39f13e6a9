// ct_parse.cpp -- CobbleTrace scene-file ingest (JSON subset + OBJ / ASCII-PLY import).
//
// Written from scratch over an in-memory buffer, but value-for-value compatible with the reference's
// ingest, because the GPU consumes exactly these doubles:
//   * numbers go through the reference's own float accumulation (fileBuffer.cpp:160-207 GetNumber):
//     digit by digit in fp32, NOT correctly rounded, then (sign*value) * pow(10, exp) rounded to float;
//   * colours are truncated double -> uint8 and packed 0x00BBGGRR (scenefile.cpp:73-74, color.h:77-80);
//   * imported vertices are placed by five row-vector x 4x4 products in the order RotY, RotX, RotZ,
//     Scale, Translate (objectLoader.cpp:27-139), each evaluated ((x*m0 + y*m1) + z*m2) + 1*m3 in fp64.
// Unlike the reference (assert / exit), malformed input raises an error.
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>
#include <thread>
#include <atomic>
#include <stdexcept>

#include "ct_scene.hpp"

namespace cth {
namespace {

struct Cursor {
    std::shared_ptr<const std::string> hold;    // the file; copies of a Cursor are cheap views of the same bytes
    const char *buf = nullptr;
    size_t size = 0;
    size_t pos = 0;
    std::string name;

    bool eof() const { return pos >= size; }
    char peek() const { return eof() ? '\0' : buf[pos]; }
    static bool is_space(char c) { return c == '\n' || c == '\r' || c == ' ' || c == '\t'; }   // fileBuffer.cpp:36-43
    void skip_space() { while (!eof() && is_space(buf[pos])) pos++; }
    void skip_line() {                                                                              // fileBuffer.cpp:52-58
        while (!eof() && buf[pos] != '\n' && buf[pos] != '\r') pos++;
        skip_space();
    }
    char token() {                                                                                  // fileBuffer.cpp:60-69
        if (eof()) return '\0';
        skip_space();
        char c = peek();
        pos++;
        return c;
    }
    void unget() { if (pos > 0) pos--; }                                                            // fileBuffer.cpp:71-78

    [[noreturn]] void fail(const std::string &what) const {
        size_t line = 1;
        for (size_t i = 0; i < pos && i < size; i++) line += buf[i] == '\n';
        std::ostringstream m;
        m << name << ":" << line << ": " << what;
        throw std::runtime_error(m.str());
    }
    void expect(char c) {                                                                           // fileBuffer.cpp:149-158
        char x = token();
        if (x != c) fail(std::string("expected '") + c + "' got '" + (x ? std::string(1, x) : std::string("EOF")) + "'");
    }

    // GetNumber, fileBuffer.cpp:160-207: fp32 accumulation; returns float.
    float number() {
        char c = token();
        float sign = 1.0f, result = 0.0f, d = 10.0f;
        bool decimal = false, scientific = false;
        int exponent = 0, esign = 1;
        if (c == '-') sign = -1.0f; else unget();
        while (!eof()) {
            c = buf[pos];
            if (c >= '0' && c <= '9') {
                if (scientific) {
                    exponent = exponent * 10 + (c - '0');
                } else if (decimal) {
                    float x = (float)(c - '0');
                    x = x / d;
                    result += x;
                    d *= 10.0f;
                } else {
                    result = result * 10.0f;
                    result += (float)(c - '0');
                }
            } else if (c == '.') {
                if (decimal) fail("second '.' in number");
                decimal = true;
            } else if (c == 'e') {
                if (scientific) fail("second 'e' in number");
                scientific = true;
            } else if (c == '-' && scientific) {
                esign = -1;
            } else {
                break;
            }
            pos++;
        }
        return (float)((double)(sign * result) * std::pow(10.0, (double)(esign * exponent)));
    }

    Vec3 vec3_bracketed() {                                                                         // GetV3 fileBuffer.cpp:209-221
        Vec3 v;
        expect('['); v.x = number(); expect(','); v.y = number(); expect(','); v.z = number(); expect(']');
        return v;
    }
    Vec3 vec3_raw() {                                                                               // GetV3Raw fileBuffer.cpp:223-234
        Vec3 v;
        skip_space(); v.x = number(); skip_space(); v.y = number(); skip_space(); v.z = number();
        return v;
    }
    std::string quoted() {                                                                          // GetString fileBuffer.cpp:80-105
        skip_space();
        if (eof()) return "";
        if (peek() != '"') fail("expected a string");
        pos++;
        std::string s;
        while (!eof() && buf[pos] != '"') s.push_back(buf[pos++]);
        pos++;
        return s;
    }
    std::string word() {                                                                            // GetStringRaw fileBuffer.cpp:107-123
        skip_space();
        std::string s;
        while (!eof() && !is_space(buf[pos])) s.push_back(buf[pos++]);
        return s;
    }
    bool boolean() {                                                                                // GetBoolean fileBuffer.cpp:125-147
        skip_space();
        std::string s;
        while (!eof() && s.size() < 5 && buf[pos] >= 'a' && buf[pos] <= 'z') s.push_back(buf[pos++]);
        if (s == "true") return true;
        if (s == "false") return false;
        fail("expected true or false");
    }
};

Cursor open_cursor(const std::string &path) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) throw std::runtime_error("cannot open " + path);
    const std::streamoff len = f.tellg();
    std::string bytes((size_t)std::max<std::streamoff>(len, 0), '\0');
    f.seekg(0);
    if (len > 0) f.read(&bytes[0], len);
    Cursor c;
    auto data = std::make_shared<const std::string>(std::move(bytes));
    c.hold = data;
    c.buf = data->data();
    c.size = data->size();
    c.name = path;
    return c;
}

// double -> uint8_t argument conversion of RgbToColor(v.x, v.y, v.z) (scenefile.cpp:74): truncation.
uint32_t pack_color(Vec3 v) {
    auto u8 = [](double x) { return (uint32_t)(uint8_t)(int)x; };
    return (u8(v.z) << 16) | (u8(v.y) << 8) | u8(v.x);
}

struct Mat4 { double m[4][4] = {{0}}; };

Mat4 identity() { Mat4 a; for (int i = 0; i < 4; i++) a.m[i][i] = 1; return a; }

// row vector (x,y,z,1) times M, keeping x,y,z (TranslatePoint objectLoader.cpp:17-26 + mymath.h:58-66)
Vec3 xform(const Mat4 &M, Vec3 p) {
    const double w = 1;
    Vec3 r;
    r.x = p.x * M.m[0][0] + p.y * M.m[1][0] + p.z * M.m[2][0] + w * M.m[3][0];
    r.y = p.x * M.m[0][1] + p.y * M.m[1][1] + p.z * M.m[2][1] + w * M.m[3][1];
    r.z = p.x * M.m[0][2] + p.y * M.m[1][2] + p.z * M.m[2][2] + w * M.m[3][2];
    return r;
}

struct Placement {            // PlaceTriangle objectLoader.cpp:27-139
    Mat4 chain[5];
    Placement(Vec3 translate, Vec3 rot, Vec3 scale) {
        Mat4 ry = identity(), rx = identity(), rz = identity(), sc = identity(), tr = identity();
        ry.m[0][0] = std::cos(rot.y); ry.m[0][2] = std::sin(rot.y); ry.m[2][0] = -std::sin(rot.y); ry.m[2][2] = std::cos(rot.y);
        rx.m[1][1] = std::cos(rot.x); rx.m[1][2] = -std::sin(rot.x); rx.m[2][1] = std::sin(rot.x); rx.m[2][2] = std::cos(rot.x);
        rz.m[0][0] = std::cos(rot.z); rz.m[0][1] = -std::sin(rot.z); rz.m[1][0] = std::sin(rot.z); rz.m[1][1] = std::cos(rot.z);
        sc.m[0][0] = scale.x; sc.m[1][1] = scale.y; sc.m[2][2] = scale.z;
        tr.m[3][0] = translate.x; tr.m[3][1] = translate.y; tr.m[3][2] = translate.z; tr.m[3][3] = 0;
        chain[0] = ry; chain[1] = rx; chain[2] = rz; chain[3] = sc; chain[4] = tr;   // application order :121-139
    }
    Triangle place(Triangle t) const {
        for (const Mat4 &M : chain) { t.p1 = xform(M, t.p1); t.p2 = xform(M, t.p2); t.p3 = xform(M, t.p3); }
        return t;
    }
};

struct ImportSpec {
    std::string filename;
    Vec3 position, rotation, scale;
    bool ply = false;          // import_format_t default IT_BLENDER (zero-initialised obj, scenefile.cpp:49)
};

std::string resolve(const std::string &base_dir, const std::string &file) {
    if (base_dir.empty() || (!file.empty() && file[0] == '/')) return file;
    return base_dir + "/" + file;
}

void push_triangle(Scene &s, const Triangle &t, const ct_material &m) {
    if (s.tris.size() >= 1000000u) throw std::runtime_error("more than MAX_OBJECTS (1000000) objects (scenefile.h:9)");
    s.tris.push_back(t);
    s.mats.push_back(m);
}

// Multi-threaded body of import_ply for well-formed files: one record per line.  The sequential
// cursor walk of the reference (parse the leading numbers of a line, skip to the next line) visits exactly the
// line starts, so the lines can be parsed independently -- with the same number() -- as long as no record runs over
// its line end.  Anything unusual (a short line, a non-triangle, an index out of range, too few lines) returns false
// and the sequential path takes over from the untouched cursor, errors included.
bool import_ply_parallel(const Cursor &c0, uint32_t n_vert, uint32_t n_face, std::vector<Vec3> &verts, const Placement &place,
                         const ct_material &mat, Scene &s) {
    const size_t n_rec = (size_t)n_vert + n_face;
    unsigned hw = std::thread::hardware_concurrency();
    int threads = (int)std::min(std::max(hw, 1u), 32u);
    if (const char *e = std::getenv("CT_HOST_THREADS")) { int v = std::atoi(e); if (v >= 1) threads = std::min(v, 64); }
    if (threads <= 1 || n_rec < 50000) return false;
    if (s.tris.size() + n_face > 1000000u) return false;            // the sequential path raises the MAX_OBJECTS error
    // where the records are: the sequential walk (take a line, skip the white space after it) visits the first non-space
    // character after every run of white space that holds a line end.  Each thread finds those in its slice of the file.
    std::vector<size_t> start, eol;
    {
        const char *buf = c0.buf;
        const size_t lo = c0.pos, size = c0.size;
        auto is_eol = [](char ch) { return ch == '\n' || ch == '\r'; };
        std::vector<std::vector<size_t>> found(threads);           // (start, eol) pairs of the records that START in the slice
        std::vector<std::thread> th;
        for (int t = 0; t < threads; t++)
            th.emplace_back([&, t] {
                const size_t a = lo + (size - lo) * (size_t)t / (size_t)threads, b = lo + (size - lo) * (size_t)(t + 1) / (size_t)threads;
                size_t p = a;
                bool between = (t == 0);                            // the header's skip_line left the cursor on the first record
                if (t > 0) {
                    size_t k = a;
                    while (k > lo && (buf[k - 1] == ' ' || buf[k - 1] == '\t')) k--;
                    between = (k > lo) && is_eol(buf[k - 1]);
                    if (k == lo) between = false;                   // still inside the first record's line
                }
                auto &out = found[t];
                out.reserve((size_t)((b - a) / 16 + 16));
                for (;;) {
                    if (!between) {
                        while (p < size && !is_eol(buf[p])) p++;
                        if (p >= b) return;                         // the next line end belongs to a later slice (or there is none)
                        between = true;
                    }
                    while (p < size && Cursor::is_space(buf[p])) p++;
                    if (p >= b || p >= size) return;                // the next record starts in a later slice
                    const size_t q = p;
                    while (p < size && !is_eol(buf[p])) p++;
                    out.push_back(q); out.push_back(p);
                    between = false;
                }
            });
        for (auto &t : th) t.join();
        size_t total = 0;
        for (auto &f : found) total += f.size() / 2;
        if (total < n_rec) return false;                            // too few lines: the sequential path reports it
        start.resize(n_rec); eol.resize(n_rec);
        size_t i = 0;
        for (auto &f : found)
            for (size_t k = 0; k + 1 < f.size() && i < n_rec; k += 2, i++) { start[i] = f[k]; eol[i] = f[k + 1]; }
    }
    std::atomic<bool> ok{true};
    auto run = [&](size_t n, auto body) {
        std::vector<std::thread> th;
        for (int t = 0; t < threads; t++)
            th.emplace_back([&, t] {
                Cursor c = c0;
                for (size_t i = n * t / threads, e = n * (t + 1) / threads; i < e && ok.load(std::memory_order_relaxed); i++)
                    if (!body(c, i)) { ok = false; return; }
            });
        for (auto &t : th) t.join();
    };
    run(n_vert, [&](Cursor &c, size_t i) {
        c.pos = start[i];
        verts[i] = c.vec3_raw();
        return c.pos <= eol[i];
    });
    if (!ok) return false;
    const size_t base = s.tris.size();
    s.tris.resize(base + n_face);
    run(n_face, [&](Cursor &c, size_t i) {
        c.pos = start[n_vert + i];
        int n = (int)c.number();
        if (n != 3) return false;
        Vec3 f = c.vec3_raw();
        if (c.pos > eol[n_vert + i]) return false;
        int a = (int)f.x, b = (int)f.y, d = (int)f.z;
        if (a < 0 || b < 0 || d < 0 || (uint32_t)a >= n_vert || (uint32_t)b >= n_vert || (uint32_t)d >= n_vert) return false;
        s.tris[base + i] = place.place(Triangle{verts[a], verts[b], verts[d]});
        return true;
    });
    if (!ok) { s.tris.resize(base); return false; }
    s.mats.resize(base + n_face, mat);
    return true;
}

// ImportPlyObject objectLoader.cpp:142-202 (ASCII PLY: x y z first on each vertex line, "3 i j k" faces)
void import_ply(const ImportSpec &spec, const ct_material &mat, const std::string &base_dir, Scene &s) {
    Cursor c = open_cursor(resolve(base_dir, spec.filename));
    if (c.word() != "ply") c.fail("not a PLY file");
    uint32_t n_vert = 0, n_face = 0;
    while (!c.eof()) {
        std::string tok = c.word();
        if (tok == "element") {
            std::string what = c.word();
            if (what == "vertex") n_vert = (uint32_t)c.number();
            else if (what == "face") n_face = (uint32_t)c.number();
        } else if (tok == "end_header") {
            c.skip_line();
            break;
        }
    }
    std::vector<Vec3> verts(n_vert);
    Placement place(spec.position, spec.rotation, spec.scale);
    const bool timing = std::getenv("CT_HOST_TIMING") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    const bool threaded = import_ply_parallel(c, n_vert, n_face, verts, place, mat, s);
    if (timing)
        fprintf(stderr, "import_ply: %u vertices, %u faces, %s: %.1f ms\n", n_vert, n_face, threaded ? "threaded" : "threaded path declined, sequential walk",
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    if (threaded) return;
    // the reference's sequential walk over the file (also the path that reports malformed input)
    for (uint32_t i = 0; i < n_vert && !c.eof(); i++) { verts[i] = c.vec3_raw(); c.skip_line(); }
    for (uint32_t i = 0; i < n_face && !c.eof(); i++) {
        int n = (int)c.number();
        if (n != 3) c.fail("only triangular faces are supported");
        Vec3 f = c.vec3_raw();
        c.skip_line();
        int a = (int)f.x, b = (int)f.y, d = (int)f.z;
        if (a < 0 || b < 0 || d < 0 || (uint32_t)a >= n_vert || (uint32_t)b >= n_vert || (uint32_t)d >= n_vert) c.fail("face index out of range");
        push_triangle(s, place.place(Triangle{verts[a], verts[b], verts[d]}), mat);
    }
}

// ImportBlenderObject objectLoader.cpp:206-279 (OBJ: 'v x y z' and 'f i j k', 1-based; '#' and 'o' lines skipped)
void import_obj(const ImportSpec &spec, const ct_material &mat, const std::string &base_dir, Scene &s) {
    Cursor c = open_cursor(resolve(base_dir, spec.filename));
    std::vector<Vec3> verts, faces;
    while (!c.eof()) {
        char t = c.token();
        if (t == '#' || t == 'o') c.skip_line();
        else if (t == 'v') { verts.push_back(c.vec3_raw()); if (verts.size() >= 1024) c.fail("more than 1023 vertices (objectLoader.cpp:222,236)"); }
        else if (t == 'f') { faces.push_back(c.vec3_raw()); if (faces.size() >= 1024) c.fail("more than 1023 faces (objectLoader.cpp:221,239)"); }
    }
    Placement place(spec.position, spec.rotation, spec.scale);
    for (const Vec3 &f : faces) {
        int a = (int)f.x - 1, b = (int)f.y - 1, d = (int)f.z - 1;
        if (a < 0 || b < 0 || d < 0 || (size_t)a >= verts.size() || (size_t)b >= verts.size() || (size_t)d >= verts.size())
            throw std::runtime_error(spec.filename + ": face index out of range");
        push_triangle(s, place.place(Triangle{verts[a], verts[b], verts[d]}), mat);
    }
}

// LoadObjects scenefile.cpp:31-131
void load_objects(Cursor &c, const std::string &base_dir, Scene &s) {
    c.expect('[');
    while (!c.eof()) {
        c.expect('{');
        enum { kSphere, kTriangle, kImport } type = kSphere;       // zero-initialised obj: OT_SPHERE
        ct_material mat{0, 0, 0.0f};
        Triangle tri;
        ImportSpec imp;
        while (!c.eof()) {
            std::string key = c.quoted();
            c.expect(':');
            if (key == "type") {
                std::string v = c.quoted();
                if (v == "sphere") type = kSphere;
                else if (v == "triangle") type = kTriangle;
                else if (v == "import") type = kImport;
                else c.fail("unknown object type '" + v + "'");
            } else if (key == "center") { (void)c.vec3_bracketed(); }
            else if (key == "radius") { (void)c.number(); }
            else if (key == "color") { mat.color = pack_color(c.vec3_bracketed()); }
            else if (key == "specular") { mat.specular = (int32_t)c.number(); }
            else if (key == "reflection") { mat.reflection = c.number(); }
            else if (key == "p1") { tri.p1 = c.vec3_bracketed(); }
            else if (key == "p2") { tri.p2 = c.vec3_bracketed(); }
            else if (key == "p3") { tri.p3 = c.vec3_bracketed(); }
            else if (key == "filename") { imp.filename = c.quoted(); }
            else if (key == "position") { imp.position = c.vec3_bracketed(); }
            else if (key == "rotation") { imp.rotation = c.vec3_bracketed(); }
            else if (key == "scale") { imp.scale = c.vec3_bracketed(); }
            else if (key == "format") {
                std::string v = c.quoted();
                if (v == "blender") imp.ply = false;
                else if (v == "ply") imp.ply = true;
            } else {
                c.fail("unknown object key '" + key + "'");
            }
            char t = c.token();
            if (t == '}') break;
            c.unget();
            c.expect(',');
        }
        if (type == kImport) {
            if (imp.ply) import_ply(imp, mat, base_dir, s); else import_obj(imp, mat, base_dir, s);
        } else if (type == kTriangle) {
            push_triangle(s, tri, mat);
        } else {
            s.n_spheres++;
        }
        char t = c.token();
        if (t == ']') break;
        c.unget();
        c.expect(',');
    }
}

// LoadLights scenefile.cpp:133-191
void load_lights(Cursor &c, Scene &s) {
    c.expect('[');
    while (!c.eof()) {
        c.expect('{');
        ct_light L{};
        L.type = CT_LIGHT_POINT;                                   // zero-initialised light_t: LT_POINT
        while (!c.eof()) {
            std::string key = c.quoted();
            c.expect(':');
            if (key == "type") {
                std::string v = c.quoted();
                if (v == "ambient") L.type = CT_LIGHT_AMBIENT;
                else if (v == "point") L.type = CT_LIGHT_POINT;
                else if (v == "directional") L.type = CT_LIGHT_DIRECTIONAL;
                else c.fail("unknown light type '" + v + "'");
            } else if (key == "intensity") { L.intensity = c.number(); }
            else if (key == "position") { Vec3 v = c.vec3_bracketed(); L.position[0] = v.x; L.position[1] = v.y; L.position[2] = v.z; }
            else if (key == "direction") { Vec3 v = c.vec3_bracketed(); L.direction[0] = v.x; L.direction[1] = v.y; L.direction[2] = v.z; }
            else c.fail("unknown light key '" + key + "'");
            char t = c.token();
            if (t == '}') break;
            c.unget();
            c.expect(',');
        }
        if (s.lights.size() >= 100u) c.fail("more than MAX_LIGHTS (100) lights (scenefile.h:10)");
        s.lights.push_back(L);
        char t = c.token();
        if (t == ']') break;
        c.unget();
        c.expect(',');
    }
}

void load_camera(Cursor &c, Scene &s) {                            // LoadCamera scenefile.cpp:193-210
    c.expect('{');
    std::string key = c.quoted();
    c.expect(':');
    if (key != "position") c.fail("camera supports only 'position'");
    s.cam_pos = c.vec3_bracketed();
    c.expect('}');
}

void load_settings(Cursor &c, Scene &s) {                          // LoadSettings scenefile.cpp:212-244
    c.expect('{');
    while (!c.eof()) {
        std::string key = c.quoted();
        c.expect(':');
        if (key == "numberOfThreads") s.settings.number_of_threads = (int32_t)c.number();
        else if (key == "subsampling") s.settings.subsampling = c.boolean();
        else if (key == "wireframe") s.settings.wireframe = c.boolean();
        else if (key == "supersampling") s.settings.supersampling = c.boolean();
        else c.fail("unknown settings key '" + key + "'");
        char t = c.token();
        if (t == '}') break;
        c.unget();
        c.expect(',');
    }
}

}  // namespace

// ParseSceneFile scenefile.cpp:256-299 (on top of InitSceneData's defaults, already in Scene{})
void parse_scene_file(const std::string &path, const std::string &base_dir, Scene &out) {
    Cursor c = open_cursor(path);
    c.expect('{');
    while (!c.eof()) {
        std::string key = c.quoted();
        c.expect(':');
        if (key == "objects") load_objects(c, base_dir, out);
        else if (key == "lights") load_lights(c, out);
        else if (key == "camera") load_camera(c, out);
        else if (key == "settings") load_settings(c, out);
        else c.fail("unknown top-level key '" + key + "'");
        char t = c.token();
        if (t == '}') break;
        c.unget();
        c.expect(',');
    }
}

// camera matrix of HandleUpdates, raythread.cpp:564-572.  yaw/pitch/roll are float there, so cos()/sin()
// resolve to the FLOAT overloads (mymath.h pulls in <math.h>) and every product/sum is fp32; only the
// store into m3x3_t widens to double.  Verified against the compiled reference (tests/test_host.py).
void camera_rotation(float yaw, float pitch, float roll, double o[9]) {
    const float cy = std::cos(yaw), sy = std::sin(yaw);
    const float cp = std::cos(pitch), sp = std::sin(pitch);
    const float cr = std::cos(roll), sr = std::sin(roll);
    o[0] = (double)(cy * cp);  o[1] = (double)(cy * sp * sr - sy * cr);  o[2] = (double)(cy * sp * cr + sy * sr);
    o[3] = (double)(sy * cp);  o[4] = (double)(sy * sp * sr + cy * cr);  o[5] = (double)(sy * sp * cr - cy * sr);
    o[6] = (double)(-sy);      o[7] = (double)(cp * sr);                 o[8] = (double)(cp * cr);
}

}  // namespace cth
