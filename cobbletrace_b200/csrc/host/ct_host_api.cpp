// ct_host_api.cpp -- extern "C" surface of include/ct_host.h.
#include <cstdio>
#include <cstring>
#include <vector>
#include <stdexcept>
#include <string>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <cerrno>
#include <thread>

#include "ct_scene.hpp"
#include "ct_tiles.hpp"

namespace cth {
struct Boss;
Boss *boss_create(Scene *scene, const ct_host_boss_config *cfg);
void boss_set_camera(Boss *b, const double pos[3], float yaw, float pitch, float roll);
void boss_reset_counter(Boss *b);
void boss_set_stream(Boss *b, int slot, void *stream);
void boss_render(Boss *b, uint32_t *bitmap, int stride, ct_host_frame_stats *stats);
int boss_tiles(const Boss *b, int32_t *out, int max_tiles);
void boss_destroy(Boss *b);
struct Controls;
Controls *controls_create(const Scene *scene, uint32_t capacity);
bool controls_add_event(Controls *c, uint32_t type, uint32_t value);
bool controls_update(Controls *c);
void controls_camera(const Controls *c, double pos[3], float ypr[3], double rot[9]);
uint32_t controls_pending(const Controls *c);
uint64_t controls_frames(const Controls *c);
void controls_destroy(Controls *c);
bool viewer_tick(Boss *b, Controls *c, uint32_t *bitmap, int stride, ct_host_present_fn present, void *user, ct_host_frame_stats *stats);
}  // namespace cth

namespace {
thread_local std::string g_err;
cth::Scene *S(ct_host_scene *s) { return reinterpret_cast<cth::Scene *>(s); }
const cth::Scene *S(const ct_host_scene *s) { return reinterpret_cast<const cth::Scene *>(s); }
}  // namespace

extern "C" {

const char *ct_host_last_error(void) { return g_err.c_str(); }

ct_host_scene *ct_host_scene_load(const char *scene_file, const char *base_dir) {
    if (!scene_file) { g_err = "scene_file is NULL"; return nullptr; }
    auto *s = new cth::Scene();
    try {
        cth::parse_scene_file(scene_file, base_dir ? base_dir : "", *s);
        if (s->tris.empty()) throw std::runtime_error(std::string(scene_file) + ": scene has no triangles (spheres are ignored by the ray tracer)");
    } catch (const std::exception &e) {
        g_err = e.what();
        delete s;
        return nullptr;
    }
    return reinterpret_cast<ct_host_scene *>(s);
}

ct_host_scene *ct_host_scene_from_arrays(uint32_t n_tri, const double *tri, const ct_material *materials, uint32_t n_lights,
                                         const ct_light *lights, const double cam_pos[3], const double cam_rot[9]) {
    if (!n_tri || !tri || !materials || (n_lights && !lights)) { g_err = "bad arrays"; return nullptr; }
    auto *s = new cth::Scene();
    s->tris.resize(n_tri);
    memcpy(static_cast<void *>(s->tris.data()), tri, (size_t)n_tri * sizeof(cth::Triangle));
    s->mats.assign(materials, materials + n_tri);
    s->lights.assign(lights, lights + n_lights);
    if (cam_pos) s->cam_pos = {cam_pos[0], cam_pos[1], cam_pos[2]};
    if (cam_rot) memcpy(s->cam_rot, cam_rot, sizeof s->cam_rot);
    return reinterpret_cast<ct_host_scene *>(s);
}

// ---- one parse + one BVH build per box: the scene in POSIX shared memory for the other processes of a multi-GPU job ----
namespace {
struct ShmHeader {
    uint64_t magic, total_bytes;
    uint32_t n_tri, n_lights, n_nodes, n_spheres;
    double cam_pos[3], cam_rot[9];
    ct_host_settings settings;
    uint64_t off_tris, off_mats, off_lights, off_nodes, off_index;
};
constexpr uint64_t kShmMagic = 0x43545343454e4531ull;      // "CTSCENE1"
size_t align64(size_t v) { return (v + 63u) & ~(size_t)63u; }
std::string shm_path(const char *name) { return std::string("/") + name; }
// big copies by a few threads (a single memcpy of 150 MB is bound by one core's page faults)
void copy_parallel(void *dst, const void *src, size_t bytes) {
    const int nt = bytes > (8u << 20) ? 4 : 1;
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++)
        th.emplace_back([=] {
            const size_t a = bytes * t / nt, b = bytes * (t + 1) / nt;
            memcpy(static_cast<char *>(dst) + a, static_cast<const char *>(src) + a, b - a);
        });
    for (auto &t : th) t.join();
}
}  // namespace

int ct_host_scene_share(const ct_host_scene *scene, const char *name) {
    if (!scene || !name || !name[0]) { g_err = "ct_host_scene_share: NULL scene or name"; return -1; }
    const cth::Scene &s = *S(scene);
    ShmHeader h{};
    h.magic = kShmMagic;
    h.n_tri = (uint32_t)s.tris.size(); h.n_lights = (uint32_t)s.lights.size(); h.n_nodes = (uint32_t)s.nodes.size(); h.n_spheres = s.n_spheres;
    h.cam_pos[0] = s.cam_pos.x; h.cam_pos[1] = s.cam_pos.y; h.cam_pos[2] = s.cam_pos.z;
    memcpy(h.cam_rot, s.cam_rot, sizeof h.cam_rot);
    h.settings = s.settings;
    size_t off = align64(sizeof h);
    h.off_tris = off; off = align64(off + s.tris.size() * sizeof(cth::Triangle));
    h.off_mats = off; off = align64(off + s.mats.size() * sizeof(ct_material));
    h.off_lights = off; off = align64(off + s.lights.size() * sizeof(ct_light));
    h.off_nodes = off; off = align64(off + s.nodes.size() * sizeof(ct_bvh_node));
    h.off_index = off; off = align64(off + s.tri_index.size() * sizeof(uint32_t));
    h.total_bytes = off;
    shm_unlink(shm_path(name).c_str());
    const int fd = shm_open(shm_path(name).c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) { g_err = std::string("shm_open(") + name + "): " + strerror(errno); return -1; }
    if (ftruncate(fd, (off_t)off) != 0) { g_err = std::string("ftruncate: ") + strerror(errno); close(fd); shm_unlink(shm_path(name).c_str()); return -1; }
    char *m = static_cast<char *>(mmap(nullptr, off, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0));
    close(fd);
    if (m == MAP_FAILED) { g_err = std::string("mmap: ") + strerror(errno); shm_unlink(shm_path(name).c_str()); return -1; }
    copy_parallel(m + h.off_tris, s.tris.data(), s.tris.size() * sizeof(cth::Triangle));
    memcpy(m + h.off_mats, s.mats.data(), s.mats.size() * sizeof(ct_material));
    memcpy(m + h.off_lights, s.lights.data(), s.lights.size() * sizeof(ct_light));
    copy_parallel(m + h.off_nodes, s.nodes.data(), s.nodes.size() * sizeof(ct_bvh_node));
    memcpy(m + h.off_index, s.tri_index.data(), s.tri_index.size() * sizeof(uint32_t));
    memcpy(m, &h, sizeof h);                             // the header last: an importer that sees the magic sees everything
    munmap(m, off);
    return 0;
}

ct_host_scene *ct_host_scene_attach(const char *name) {
    if (!name || !name[0]) { g_err = "ct_host_scene_attach: NULL name"; return nullptr; }
    const int fd = shm_open(shm_path(name).c_str(), O_RDONLY, 0);
    if (fd < 0) { g_err = std::string("shm_open(") + name + "): " + strerror(errno); return nullptr; }
    struct stat st;
    if (fstat(fd, &st) != 0 || (size_t)st.st_size < sizeof(ShmHeader)) { g_err = "shared scene is truncated"; close(fd); return nullptr; }
    const char *m = static_cast<const char *>(mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_SHARED, fd, 0));
    close(fd);
    if (m == MAP_FAILED) { g_err = std::string("mmap: ") + strerror(errno); return nullptr; }
    ShmHeader h;
    memcpy(&h, m, sizeof h);
    if (h.magic != kShmMagic || h.total_bytes > (uint64_t)st.st_size) { g_err = "shared scene has a bad header"; munmap(const_cast<char *>(m), (size_t)st.st_size); return nullptr; }
    auto *s = new cth::Scene();
    s->tris.resize(h.n_tri); s->mats.resize(h.n_tri); s->lights.resize(h.n_lights); s->nodes.resize(h.n_nodes); s->tri_index.resize(h.n_nodes ? h.n_tri : 0);
    copy_parallel(static_cast<void *>(s->tris.data()), m + h.off_tris, (size_t)h.n_tri * sizeof(cth::Triangle));
    memcpy(s->mats.data(), m + h.off_mats, (size_t)h.n_tri * sizeof(ct_material));
    memcpy(s->lights.data(), m + h.off_lights, (size_t)h.n_lights * sizeof(ct_light));
    copy_parallel(static_cast<void *>(s->nodes.data()), m + h.off_nodes, (size_t)h.n_nodes * sizeof(ct_bvh_node));
    memcpy(s->tri_index.data(), m + h.off_index, s->tri_index.size() * sizeof(uint32_t));
    s->cam_pos = {h.cam_pos[0], h.cam_pos[1], h.cam_pos[2]};
    memcpy(s->cam_rot, h.cam_rot, sizeof h.cam_rot);
    s->settings = h.settings; s->n_spheres = h.n_spheres;
    munmap(const_cast<char *>(m), (size_t)st.st_size);
    return reinterpret_cast<ct_host_scene *>(s);
}

int ct_host_scene_unshare(const char *name) {
    if (!name || !name[0]) return -1;
    return shm_unlink(shm_path(name).c_str()) == 0 ? 0 : -1;
}

void ct_host_scene_free(ct_host_scene *s) { delete S(s); }

uint32_t ct_host_scene_triangle_count(const ct_host_scene *s) { return (uint32_t)S(s)->tris.size(); }
uint32_t ct_host_scene_sphere_count(const ct_host_scene *s) { return S(s)->n_spheres; }
const double *ct_host_scene_triangles(const ct_host_scene *s) { return &S(s)->tris.data()->p1.x; }
const ct_material *ct_host_scene_materials(const ct_host_scene *s) { return S(s)->mats.data(); }
uint32_t ct_host_scene_light_count(const ct_host_scene *s) { return (uint32_t)S(s)->lights.size(); }
const ct_light *ct_host_scene_lights(const ct_host_scene *s) { return S(s)->lights.data(); }

void ct_host_scene_camera(const ct_host_scene *s, double pos[3], double rot[9]) {
    if (pos) { pos[0] = S(s)->cam_pos.x; pos[1] = S(s)->cam_pos.y; pos[2] = S(s)->cam_pos.z; }
    if (rot) memcpy(rot, S(s)->cam_rot, sizeof S(s)->cam_rot);
}

void ct_host_scene_set_camera(ct_host_scene *s, const double pos[3], const double rot[9]) {
    if (pos) S(s)->cam_pos = {pos[0], pos[1], pos[2]};
    if (rot) memcpy(S(s)->cam_rot, rot, sizeof S(s)->cam_rot);
}

void ct_host_scene_settings(const ct_host_scene *s, ct_host_settings *out) { if (out) *out = S(s)->settings; }

void ct_host_scene_set_reflection(ct_host_scene *s, float reflection) {
    for (auto &m : S(s)->mats) m.reflection = reflection;
}

int ct_host_scene_render_flags(const ct_host_scene *s, uint32_t *flags) {
    if (!s || !flags) { g_err = "NULL scene or flags"; return CT_ERR_INVALID; }
    const ct_host_settings &st = S(s)->settings;
    *flags = (st.subsampling ? (uint32_t)CT_FLAG_SUBSAMPLING : 0u) | (st.supersampling ? (uint32_t)CT_FLAG_SUPERSAMPLING : 0u);
    return CT_OK;
}

int ct_host_build_bvh(ct_host_scene *s) {
    try {
        if (S(s)->nodes.empty()) cth::build_bvh(*S(s));
    } catch (const std::exception &e) {
        g_err = e.what();
        return CT_ERR_INVALID;
    }
    return (int)S(s)->nodes.size();
}

const ct_bvh_node *ct_host_scene_nodes(const ct_host_scene *s, uint32_t *n_nodes) {
    if (n_nodes) *n_nodes = (uint32_t)S(s)->nodes.size();
    return S(s)->nodes.data();
}

const uint32_t *ct_host_scene_tri_indexes(const ct_host_scene *s) { return S(s)->tri_index.data(); }

int ct_host_scene_set_bvh(ct_host_scene *s, uint32_t n_nodes, const ct_bvh_node *nodes, const uint32_t *tri_indexes) {
    if (!n_nodes || !nodes || !tri_indexes) { g_err = "bad BVH arrays"; return CT_ERR_INVALID; }
    S(s)->nodes.assign(nodes, nodes + n_nodes);
    S(s)->tri_index.assign(tri_indexes, tri_indexes + S(s)->tris.size());
    return CT_OK;
}

void ct_host_camera_rotation(float yaw, float pitch, float roll, double out[9]) { cth::camera_rotation(yaw, pitch, roll, out); }

int ct_host_fill_desc(const ct_host_scene *s, int width, int height, int max_depth, uint32_t flags, ct_scene_desc *d) {
    const cth::Scene *sc = S(s);
    if (!d) { g_err = "NULL desc"; return CT_ERR_INVALID; }
    if (sc->nodes.empty()) { g_err = "scene has no BVH yet: call ct_host_build_bvh first"; return CT_ERR_INVALID; }
    memset(d, 0, sizeof *d);
    d->struct_size = sizeof *d;
    d->flags = flags;
    d->n_triangles = (uint32_t)sc->tris.size();
    d->triangle_stride = sizeof(cth::Triangle);
    d->triangles = sc->tris.data();
    d->materials = sc->mats.data();
    d->n_nodes = (uint32_t)sc->nodes.size();
    d->n_lights = (uint32_t)sc->lights.size();
    d->nodes = sc->nodes.data();
    d->tri_indexes = sc->tri_index.data();
    d->lights = sc->lights.data();
    d->camera_position[0] = sc->cam_pos.x; d->camera_position[1] = sc->cam_pos.y; d->camera_position[2] = sc->cam_pos.z;
    memcpy(d->camera_rotation, sc->cam_rot, sizeof d->camera_rotation);
    d->viewport[0] = d->viewport[1] = d->viewport[2] = 1.0f;       // raythread.cpp:554
    d->width = width; d->height = height; d->max_depth = max_depth;
    d->background = 0x333333u;                                     // raythread.cpp:59
    return CT_OK;
}

ct_host_boss *ct_host_boss_create(ct_host_scene *s, const ct_host_boss_config *cfg) {
    try {
        return reinterpret_cast<ct_host_boss *>(cth::boss_create(S(s), cfg));
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}

#define GUARD(stmt) try { stmt; } catch (const std::exception &e) { g_err = e.what(); return CT_ERR_CUDA; } return CT_OK

int ct_host_boss_set_camera(ct_host_boss *b, const double pos[3], float yaw, float pitch, float roll) {
    GUARD(cth::boss_set_camera(reinterpret_cast<cth::Boss *>(b), pos, yaw, pitch, roll));
}

int ct_host_boss_set_stream(ct_host_boss *b, int device_slot, void *cuda_stream) {
    GUARD(cth::boss_set_stream(reinterpret_cast<cth::Boss *>(b), device_slot, cuda_stream));
}

int ct_host_boss_reset_shared_counter(ct_host_boss *b) { GUARD(cth::boss_reset_counter(reinterpret_cast<cth::Boss *>(b))); }

int ct_host_boss_render(ct_host_boss *b, uint32_t *bitmap, int stride_pixels, ct_host_frame_stats *stats) {
    GUARD(cth::boss_render(reinterpret_cast<cth::Boss *>(b), bitmap, stride_pixels, stats));
}

int ct_host_boss_tiles(const ct_host_boss *b, int32_t *y_ranges, int max_tiles) {
    return cth::boss_tiles(reinterpret_cast<const cth::Boss *>(b), y_ranges, max_tiles);
}

int ct_host_write_ppm(const char *path, const uint32_t *bitmap, int width, int height, int stride_pixels) {
    if (!path || !bitmap || width <= 0 || height <= 0 || stride_pixels < width) { g_err = "ct_host_write_ppm: bad arguments"; return CT_ERR_INVALID; }
    FILE *f = fopen(path, "wb");
    if (!f) { g_err = std::string("cannot open ") + path; return CT_ERR_INVALID; }
    fprintf(f, "P6\n%d %d\n255\n", width, height);
    std::vector<unsigned char> row((size_t)width * 3);
    bool ok = true;
    for (int y = 0; y < height && ok; y++) {
        const uint32_t *src = bitmap + (size_t)y * stride_pixels;
        for (int x = 0; x < width; x++) {
            row[3 * (size_t)x + 0] = (unsigned char)(src[x] & 0xFF);             // red is the low byte
            row[3 * (size_t)x + 1] = (unsigned char)((src[x] >> 8) & 0xFF);
            row[3 * (size_t)x + 2] = (unsigned char)((src[x] >> 16) & 0xFF);
        }
        ok = fwrite(row.data(), 1, row.size(), f) == row.size();
    }
    ok = (fclose(f) == 0) && ok;
    if (!ok) { g_err = std::string("short write to ") + path; return CT_ERR_INVALID; }
    return CT_OK;
}

ct_host_tile_counter *ct_host_tile_counter_open(const char *shared_name) {
    auto *c = new cth::TileCounter();
    try {
        if (shared_name && shared_name[0]) c->open_shared(shared_name);
    } catch (const std::exception &e) {
        g_err = e.what();
        delete c;
        return nullptr;
    }
    return reinterpret_cast<ct_host_tile_counter *>(c);
}
int32_t ct_host_tile_counter_next(ct_host_tile_counter *c) { return reinterpret_cast<cth::TileCounter *>(c)->next(); }
void ct_host_tile_counter_reset(ct_host_tile_counter *c) { reinterpret_cast<cth::TileCounter *>(c)->reset(); }
void ct_host_tile_counter_close(ct_host_tile_counter *c, int unlink_shared) {
    auto *t = reinterpret_cast<cth::TileCounter *>(c);
    if (t && unlink_shared && !t->shm_name.empty()) shm_unlink(t->shm_name.c_str());
    delete t;
}

void ct_host_boss_destroy(ct_host_boss *b) { cth::boss_destroy(reinterpret_cast<cth::Boss *>(b)); }

#define CTL(c) reinterpret_cast<cth::Controls *>(c)
#define CCTL(c) reinterpret_cast<const cth::Controls *>(c)

ct_host_controls *ct_host_controls_create(const ct_host_scene *s, uint32_t event_capacity) {
    if (!s) { g_err = "NULL scene"; return nullptr; }
    return reinterpret_cast<ct_host_controls *>(cth::controls_create(S(s), event_capacity));
}

int ct_host_controls_add_event(ct_host_controls *c, uint32_t type, uint32_t value) {
    if (!c) { g_err = "NULL controls"; return CT_ERR_INVALID; }
    if (!cth::controls_add_event(CTL(c), type, value)) { g_err = "event queue full"; return CT_ERR_INVALID; }
    return CT_OK;
}

int ct_host_controls_update(ct_host_controls *c) {
    if (!c) { g_err = "NULL controls"; return CT_ERR_INVALID; }
    return cth::controls_update(CTL(c)) ? 1 : 0;
}

uint32_t ct_host_controls_pending(const ct_host_controls *c) { return c ? cth::controls_pending(CCTL(c)) : 0; }
uint64_t ct_host_controls_frames(const ct_host_controls *c) { return c ? cth::controls_frames(CCTL(c)) : 0; }

void ct_host_controls_camera(const ct_host_controls *c, double pos[3], float ypr[3], double rot[9]) {
    if (c) cth::controls_camera(CCTL(c), pos, ypr, rot);
}

void ct_host_controls_destroy(ct_host_controls *c) { cth::controls_destroy(CTL(c)); }

int ct_host_viewer_tick(ct_host_boss *b, ct_host_controls *c, uint32_t *bitmap, int stride_pixels, ct_host_present_fn present,
                        void *user, int *rendered, ct_host_frame_stats *stats) {
    if (!b || !c || !bitmap) { g_err = "ct_host_viewer_tick: NULL boss, controls or bitmap"; return CT_ERR_INVALID; }
    try {
        const bool r = cth::viewer_tick(reinterpret_cast<cth::Boss *>(b), CTL(c), bitmap, stride_pixels, present, user, stats);
        if (rendered) *rendered = r ? 1 : 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return CT_ERR_CUDA;
    }
    return CT_OK;
}

}  // extern "C"
