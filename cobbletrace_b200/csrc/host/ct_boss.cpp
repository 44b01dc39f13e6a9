// ct_boss.cpp -- the boss half of RayThread (raythread.cpp:641-666 + HandleUpdates :546-594), with the
// SDL worker threads replaced by ct_gpu_render_tile calls.
//
// Differences from the reference, all deliberate:
//   * the static split `yStep = H / numberOfThreads` (:576, loses H % N rows) becomes dynamic tile
//     stealing: row tiles are handed out by an atomic counter -- process-local for the GPUs this process
//     drives (one host thread per GPU), or a POSIX shared-memory counter when one process per GPU is
//     used (torchrun): load balance follows the image content, no row is ever dropped;
//   * completion is a stream synchronisation, not polling of `status` flags (:657-661);
//   * with several GPUs in one process and no tile size given, the frame is ONE shared tile: the devices' own
//     warps trace 32-pixel chunks (most dealt round-robin, the rest stolen from a cursor on devices[0] over NVLink:
//     ct_gpu_share_partition) and store finished pixels straight into its framebuffer (ct_gpu_render_shared); with a tile size, finished row tiles of the other GPUs are copied
//     to devices[0] (ct_gpu_gather_rows -> cudaMemcpyPeerAsync) before the single readback.
// libct_gpu.so is loaded with dlopen so this library also loads on machines without CUDA; rendering
// then fails loudly (there is no CPU path).
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "ct_scene.hpp"
#include "ct_tiles.hpp"

namespace cth {

struct GpuApi {
    void *handle = nullptr;
    int (*abi_version)() = nullptr;
    const char *(*last_error)() = nullptr;
    int (*upload_scene)(int, const ct_scene_desc *) = nullptr;
    int (*set_camera)(int, const double *, const double *) = nullptr;
    int (*set_stream)(int, void *) = nullptr;
    int (*kernel_launches)(int, uint64_t *, int) = nullptr;
    int (*render_tile)(int, int, int, ct_ray_counters *) = nullptr;
    int (*throttle)(int, int) = nullptr;
    int (*readback)(int, uint32_t *, int, int, int) = nullptr;
    int (*get_counters)(int, ct_ray_counters *, int) = nullptr;
    int (*sync)(int) = nullptr;
    int (*last_tile_ms)(int, float *) = nullptr;
    int (*gather_rows)(int, int, int, int) = nullptr;
    int (*share_export)(int, ct_gpu_share *) = nullptr;
    int (*share_attach)(int, const ct_gpu_share *) = nullptr;
    int (*share_partition)(int, int, int) = nullptr;
    int (*share_reset)(int) = nullptr;
    int (*render_shared)(int, int, int, ct_ray_counters *) = nullptr;
    int (*shutdown)(int) = nullptr;

    void load(const std::string &path) {
        handle = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!handle) throw std::runtime_error(std::string("cannot load the CUDA renderer (no CPU fallback): ") + dlerror());
        auto sym = [&](const char *name) {
            void *p = dlsym(handle, name);
            if (!p) throw std::runtime_error(std::string("libct_gpu.so lacks symbol ") + name);
            return p;
        };
        abi_version = (int (*)())sym("ct_gpu_abi_version");
        last_error = (const char *(*)())sym("ct_gpu_last_error");
        upload_scene = (int (*)(int, const ct_scene_desc *))sym("ct_gpu_upload_scene");
        set_camera = (int (*)(int, const double *, const double *))sym("ct_gpu_set_camera");
        set_stream = (int (*)(int, void *))sym("ct_gpu_set_stream");
        kernel_launches = (int (*)(int, uint64_t *, int))sym("ct_gpu_kernel_launches");
        render_tile = (int (*)(int, int, int, ct_ray_counters *))sym("ct_gpu_render_tile");
        throttle = (int (*)(int, int))sym("ct_gpu_throttle");
        readback = (int (*)(int, uint32_t *, int, int, int))sym("ct_gpu_readback");
        get_counters = (int (*)(int, ct_ray_counters *, int))sym("ct_gpu_get_counters");
        sync = (int (*)(int))sym("ct_gpu_sync");
        last_tile_ms = (int (*)(int, float *))sym("ct_gpu_last_tile_ms");
        gather_rows = (int (*)(int, int, int, int))sym("ct_gpu_gather_rows");
        share_export = (int (*)(int, ct_gpu_share *))sym("ct_gpu_share_export");
        share_attach = (int (*)(int, const ct_gpu_share *))sym("ct_gpu_share_attach");
        share_partition = (int (*)(int, int, int))sym("ct_gpu_share_partition");
        share_reset = (int (*)(int))sym("ct_gpu_share_reset");
        render_shared = (int (*)(int, int, int, ct_ray_counters *))sym("ct_gpu_render_shared");
        shutdown = (int (*)(int))sym("ct_gpu_shutdown");
        if (abi_version() != CT_GPU_ABI_VERSION) throw std::runtime_error("libct_gpu.so ABI version mismatch");
    }
    void check(int rc, const char *what) const {
        if (rc < 0) throw std::runtime_error(std::string(what) + ": " + last_error());
    }
};

struct Boss {
    Scene *scene = nullptr;
    ct_host_boss_config cfg{};
    GpuApi gpu;
    TileCounter counter;
    int tile_rows = 0, n_tiles = 0, y_lo = 0, y_hi = 0;
    bool shared_frame = false;      // several GPUs of this process render ONE tile, stealing chunks on the device (ct_gpu_render_shared)
    std::vector<std::vector<std::pair<int, int>>> tiles_by_dev;   // last frame
    std::vector<double> device_ms;  // last frame: kernel time of the tiles each device rendered
    std::string error;
};

static std::string default_gpu_library() {
    Dl_info info;
    if (dladdr((void *)&default_gpu_library, &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        size_t k = p.find_last_of('/');
        return (k == std::string::npos ? std::string(".") : p.substr(0, k)) + "/libct_gpu.so";
    }
    return "libct_gpu.so";
}

Boss *boss_create(Scene *scene, const ct_host_boss_config *cfg) {
    if (!scene || !cfg || cfg->struct_size != sizeof(ct_host_boss_config)) throw std::runtime_error("bad boss config (struct_size)");
    if (cfg->n_devices < 1 || cfg->n_devices > 16) throw std::runtime_error("n_devices must be 1..16");
    if (cfg->width <= 0 || cfg->height <= 0) throw std::runtime_error("bad bitmap size");
    auto *b = new Boss();
    try {
        b->scene = scene;
        b->cfg = *cfg;
        // settings.subsampling stores, for every traced row, an averaged pixel into the row BELOW it (raythread.cpp:512-531): a
        // tile writes one row of its neighbour.  Tiles of one device are rendered in order and resolve that seam like the
        // reference's sequential loop; tiles on different devices would race on it (and the gather copies only a tile's own
        // rows), so the frame must stay on one device.
        if ((cfg->flags & CT_FLAG_SUBSAMPLING) &&
            (cfg->n_devices > 1 || (cfg->shared_counter_name && cfg->shared_counter_name[0] && cfg->world_size > 1)))
            throw std::runtime_error("CT_FLAG_SUBSAMPLING needs the whole frame on one device (a tile writes into the row below it); use one device");
        b->gpu.load(cfg->gpu_library && cfg->gpu_library[0] ? cfg->gpu_library : default_gpu_library());
        if (scene->nodes.empty()) build_bvh(*scene);                               // raythread.cpp:650-651
        ct_scene_desc d;
        if (ct_host_fill_desc(reinterpret_cast<ct_host_scene *>(scene), cfg->width, cfg->height, cfg->max_depth, cfg->flags, &d) != 0)
            throw std::runtime_error(ct_host_last_error());
        for (int i = 0; i < cfg->n_devices; i++) b->gpu.check(b->gpu.upload_scene(cfg->devices[i], &d), "ct_gpu_upload_scene");
        const bool shared = cfg->shared_counter_name && cfg->shared_counter_name[0];
        if (shared) b->counter.open_shared(cfg->shared_counter_name);
        // canvas rows the reference's loops cover when the thread count divides H (raythread.cpp:574-581):
        // y in [-(H/2), -(H/2) + H)
        const int half = cfg->height / 2;
        b->y_lo = -half; b->y_hi = -half + cfg->height;
        const int total_devices = shared ? std::max(1, cfg->world_size) * cfg->n_devices : cfg->n_devices;
        int rows = cfg->tile_rows;
        if (rows <= 0) {
            if (total_devices == 1) rows = b->y_hi - b->y_lo;                        // one tile: whole frame in flight
            else rows = std::max(8, ((b->y_hi - b->y_lo) / (total_devices * 8) + 3) / 4 * 4);
        }
        // Several GPUs in this process and no tile size asked for: one shared frame -- the devices' own warps take
        // 32-pixel chunks (dealt, and stolen from a cursor on devices[0] over NVLink) and store their pixels into its framebuffer.
        b->shared_frame = !shared && cfg->n_devices > 1 && cfg->tile_rows <= 0 && !(cfg->flags & CT_FLAG_SUBSAMPLING);
        if (b->shared_frame) {
            rows = b->y_hi - b->y_lo;
            ct_gpu_share h;
            memset(&h, 0, sizeof h);
            h.struct_size = sizeof h;
            b->gpu.check(b->gpu.share_export(cfg->devices[0], &h), "ct_gpu_share_export");
            for (int i = 1; i < cfg->n_devices; i++) b->gpu.check(b->gpu.share_attach(cfg->devices[i], &h), "ct_gpu_share_attach");
            // every device renders every frame: deal most chunks round-robin, steal the rest (ct_gpu_share_partition)
            for (int i = 0; i < cfg->n_devices; i++) b->gpu.check(b->gpu.share_partition(cfg->devices[i], i, cfg->n_devices), "ct_gpu_share_partition");
        }
        b->tile_rows = rows;
        b->n_tiles = (b->y_hi - b->y_lo + rows - 1) / rows;
        b->tiles_by_dev.resize(cfg->n_devices);
        b->device_ms.assign(cfg->n_devices, 0.0);
    } catch (...) {
        delete b;
        throw;
    }
    return b;
}

void boss_set_camera(Boss *b, const double pos[3], float yaw, float pitch, float roll) {
    double rot[9];
    camera_rotation(yaw, pitch, roll, rot);
    b->scene->cam_pos = {pos[0], pos[1], pos[2]};
    memcpy(b->scene->cam_rot, rot, sizeof rot);
    for (int i = 0; i < b->cfg.n_devices; i++) b->gpu.check(b->gpu.set_camera(b->cfg.devices[i], pos, rot), "ct_gpu_set_camera");
}

void boss_reset_counter(Boss *b) { b->counter.reset(); }

void boss_set_stream(Boss *b, int slot, void *stream) {
    if (slot < 0 || slot >= b->cfg.n_devices) throw std::runtime_error("device slot out of range");
    b->gpu.check(b->gpu.set_stream(b->cfg.devices[slot], stream), "ct_gpu_set_stream");
}

void boss_render(Boss *b, uint32_t *bitmap, int stride, ct_host_frame_stats *stats) {
    const auto t0 = std::chrono::steady_clock::now();
    const int nd = b->cfg.n_devices;
    const bool shared = b->counter.shared != nullptr;
    if (!shared) b->counter.reset();
    for (auto &v : b->tiles_by_dev) v.clear();
    std::fill(b->device_ms.begin(), b->device_ms.end(), 0.0);
    std::vector<std::string> errs(nd);
    auto worker = [&](int k) {
        try {
            const int dev = b->cfg.devices[k];
            if (b->shared_frame) {                       // every device renders the same tile; the split happens on the devices
                b->gpu.check(b->gpu.render_shared(dev, b->y_lo, b->y_hi, nullptr), "ct_gpu_render_shared");
                b->tiles_by_dev[k].push_back({b->y_lo, b->y_hi});
                b->gpu.check(b->gpu.sync(dev), "ct_gpu_sync");
                float ms = 0.0f;
                if (b->gpu.last_tile_ms(dev, &ms) >= 0) b->device_ms[k] += ms;
                return;
            }
            while (true) {
                int t = b->counter.next();
                if (t >= b->n_tiles) break;
                int y0 = b->y_lo + t * b->tile_rows, y1 = std::min(b->y_hi, y0 + b->tile_rows);
                b->gpu.check(b->gpu.render_tile(dev, y0, y1, nullptr), "ct_gpu_render_tile");
                b->tiles_by_dev[k].push_back({y0, y1});
                // keep at most two tiles in flight per GPU so that stealing follows real progress
                b->gpu.check(b->gpu.throttle(dev, 1), "ct_gpu_throttle");
                if (stats && b->n_tiles > 1) {           // per-tile kernel times are only read when somebody asks (it waits for the tile)
                    float ms = 0.0f;
                    if (b->gpu.last_tile_ms(dev, &ms) >= 0) b->device_ms[k] += ms;
                }
            }
            // finished tiles -> devices[0] over NVLink
            if (k != 0) {
                const int H = b->cfg.height, half = H / 2;
                for (auto [y0, y1] : b->tiles_by_dev[k])
                    b->gpu.check(b->gpu.gather_rows(dev, b->cfg.devices[0], half - (y1 - 1), half - y0 + 1), "ct_gpu_gather_rows");
            }
            b->gpu.check(b->gpu.sync(dev), "ct_gpu_sync");
            if (b->n_tiles == 1 && !b->tiles_by_dev[k].empty()) {
                float ms = 0.0f;
                if (b->gpu.last_tile_ms(dev, &ms) >= 0) b->device_ms[k] += ms;
            }
        } catch (const std::exception &e) {
            errs[k] = e.what();
        }
    };
    if (nd == 1) {
        worker(0);
    } else {
        std::vector<std::thread> th;
        for (int k = 0; k < nd; k++) th.emplace_back(worker, k);
        for (auto &t : th) t.join();
    }
    for (auto &e : errs) if (!e.empty()) throw std::runtime_error(e);
    if (bitmap) {
        const int H = b->cfg.height, half = H / 2;
        if (!shared) {
            b->gpu.check(b->gpu.readback(b->cfg.devices[0], bitmap, stride, 0, H), "ct_gpu_readback");
        } else {
            for (int k = 0; k < nd; k++)
                for (auto [y0, y1] : b->tiles_by_dev[k])
                    b->gpu.check(b->gpu.readback(b->cfg.devices[0], bitmap, stride, half - (y1 - 1), half - y0 + 1), "ct_gpu_readback");
        }
    }
    if (stats) {
        memset(stats, 0, sizeof *stats);
        for (int k = 0; k < nd; k++) {
            ct_ray_counters c;
            b->gpu.check(b->gpu.get_counters(b->cfg.devices[k], &c, 1), "ct_gpu_get_counters");
            stats->rays.rays_primary += c.rays_primary; stats->rays.rays_shadow += c.rays_shadow;
            stats->rays.rays_reflection += c.rays_reflection; stats->rays.box_tests += c.box_tests; stats->rays.tri_tests += c.tri_tests;
            stats->tiles_mine += (int)b->tiles_by_dev[k].size();
            uint64_t nl = 0;
            b->gpu.check(b->gpu.kernel_launches(b->cfg.devices[k], &nl, 1), "ct_gpu_kernel_launches");
            stats->kernel_launches += nl;
        }
        stats->tiles_total = b->n_tiles;
        stats->device_ms_max = (float)*std::max_element(b->device_ms.begin(), b->device_ms.end());
        stats->wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
}

int boss_tiles(const Boss *b, int32_t *out, int max_tiles) {
    int n = 0;
    for (auto &v : b->tiles_by_dev)
        for (auto [y0, y1] : v) {
            if (n < max_tiles && out) { out[2 * n] = y0; out[2 * n + 1] = y1; }
            n++;
        }
    return n;
}

void boss_destroy(Boss *b) {
    if (!b) return;
    // nothing may still be writing into devices[0] (peer stores, gathered rows) when its memory goes: drain every
    // device first, then let go of the attached ones before the root
    if (b->gpu.sync) for (int i = 0; i < b->cfg.n_devices; i++) b->gpu.sync(b->cfg.devices[i]);
    if (b->gpu.shutdown) for (int i = b->cfg.n_devices - 1; i >= 0; i--) b->gpu.shutdown(b->cfg.devices[i]);
    delete b;
}

}  // namespace cth
