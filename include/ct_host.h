/* include/ct_host.h -- C ABI of the HOST side that feeds include/ct_gpu.h.
 *
 * The reference keeps these steps in C++ on the host and so does this repo (libct_host.so, g++ only,
 * loads without CUDA).  They mirror, bit for bit, what the GPU renderer consumes:
 *
 *   ct_host_scene_load      InitSceneData + ParseSceneFile + ImportObject   scenefile.cpp:12,257 / objectLoader.cpp:282
 *   ct_host_build_bvh       GetSceneTriangles + InitializeBVHState + BuildBVH  raythread.cpp:621, bvh.cpp:16,108
 *   ct_host_camera_rotation the camera matrix of HandleUpdates               raythread.cpp:564-572
 *   ct_host_boss_*          RayThread's boss half: dispatch row tiles, wait, hand back the bitmap
 *                           (raythread.cpp:641-666, 546-594) -- workers are ct_gpu_render_tile / ct_gpu_render_shared
 *                           calls on one or more GPUs, with dynamic stealing instead of the static yStep split.
 *   ct_host_controls_*      the event queue + HandleKeyboard + HandleUpdates' camera state  eventQueue.cpp, raythread.cpp:388-434,546-572
 *   ct_host_viewer_tick     one main-loop iteration (RayThread + Blit seam)     cobbletrace.cpp:88-118
 *
 * All functions are thread-compatible (one thread per object).  Errors: NULL / negative return and a
 * message via ct_host_last_error(); nothing asserts or exits (the reference's parser asserts).
 */
#ifndef CT_HOST_H
#define CT_HOST_H

#include <stddef.h>
#include <stdint.h>

#include "ct_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ct_host_scene ct_host_scene;
typedef struct ct_host_boss ct_host_boss;
typedef struct ct_host_tile_counter ct_host_tile_counter;

/* == settings_t, reference scenefile.h:83-88 (defaults from InitSceneData scenefile.cpp:23-26) */
typedef struct ct_host_settings {
    int32_t number_of_threads;
    int32_t subsampling, wireframe, supersampling;
} ct_host_settings;

const char *ct_host_last_error(void);

/* ---- scene ingest ------------------------------------------------------------------------------------ */
/* Parse a CobbleTrace scene file (JSON subset) including OBJ ("blender") / ASCII-PLY imports.  Relative
 * model paths resolve against the process cwd, exactly like the reference (fopen in fileBuffer.cpp:237);
 * pass base_dir != NULL to resolve them against a directory instead. */
ct_host_scene *ct_host_scene_load(const char *scene_file, const char *base_dir);

/* Build a scene from flattened arrays (tri: n_tri x 9 doubles). Copies everything. */
ct_host_scene *ct_host_scene_from_arrays(uint32_t n_tri, const double *tri, const ct_material *materials,
                                         uint32_t n_lights, const ct_light *lights,
                                         const double cam_pos[3], const double cam_rot[9]);
void ct_host_scene_free(ct_host_scene *s);

/* One parse and one BVH build per box.  In a job with one process per GPU the scene is replicated on every GPU but need
 * not be parsed and built by every process: one of them calls ct_host_scene_share (scene + BVH as built go into the POSIX
 * shared-memory object `name`), the others ct_host_scene_attach (a private copy of those arrays -- a memcpy instead of
 * ParseSceneFile + BuildBVH), and the first one ct_host_scene_unshare once everybody has attached.  0 / non-NULL on success. */
int ct_host_scene_share(const ct_host_scene *s, const char *name);
ct_host_scene *ct_host_scene_attach(const char *name);
int ct_host_scene_unshare(const char *name);

uint32_t ct_host_scene_triangle_count(const ct_host_scene *s);
uint32_t ct_host_scene_sphere_count(const ct_host_scene *s);      /* parsed but ignored by the tracer (raythread.cpp:208) */
const double *ct_host_scene_triangles(const ct_host_scene *s);    /* n_tri x 9, GetSceneTriangles order */
const ct_material *ct_host_scene_materials(const ct_host_scene *s);
uint32_t ct_host_scene_light_count(const ct_host_scene *s);
const ct_light *ct_host_scene_lights(const ct_host_scene *s);
void ct_host_scene_camera(const ct_host_scene *s, double pos[3], double rot[9]);
void ct_host_scene_set_camera(ct_host_scene *s, const double pos[3], const double rot[9]);
void ct_host_scene_settings(const ct_host_scene *s, ct_host_settings *out);
/* The CT_FLAG_* a renderer needs to honour the scene file's own settings (raythread.cpp:460,512): subsampling ->
 * CT_FLAG_SUBSAMPLING, supersampling -> CT_FLAG_SUPERSAMPLING (deterministic jitter, see ct_gpu.h); they combine as in the
 * reference.  Benchmarks and parity tests pass 0 instead (sampling off). */
int ct_host_scene_render_flags(const ct_host_scene *s, uint32_t *flags);
/* Harness-level material override (BASELINE config 3: "reflection > 0 forced"). */
void ct_host_scene_set_reflection(ct_host_scene *s, float reflection);

/* ---- BVH ----------------------------------------------------------------------------------------------- */
/* Midpoint-split BVH, node numbering and triangle permutation identical to the reference's BuildBVH.
 * Returns nodesUsed (>0) or a negative ct_status. Idempotent. */
int ct_host_build_bvh(ct_host_scene *s);
const ct_bvh_node *ct_host_scene_nodes(const ct_host_scene *s, uint32_t *n_nodes);
const uint32_t *ct_host_scene_tri_indexes(const ct_host_scene *s);
/* Install a BVH built elsewhere (e.g. by the reference's own bvh.cpp). Copies. */
int ct_host_scene_set_bvh(ct_host_scene *s, uint32_t n_nodes, const ct_bvh_node *nodes, const uint32_t *tri_indexes);

void ct_host_camera_rotation(float yaw, float pitch, float roll, double out[9]);

/* Fill a ct_scene_desc that borrows the scene's arrays (valid while the scene lives). */
int ct_host_fill_desc(const ct_host_scene *s, int width, int height, int max_depth, uint32_t flags, ct_scene_desc *out);

/* ---- boss ------------------------------------------------------------------------------------------------ */
typedef struct ct_host_boss_config {
    uint32_t struct_size;
    int32_t width, height;        /* bitmapSettings_t */
    int32_t max_depth;            /* 10 = reference */
    uint32_t flags;               /* CT_FLAG_* */
    int32_t n_devices;            /* GPUs driven by THIS process (one host thread each) */
    int32_t devices[16];
    int32_t tile_rows;            /* rows per stolen tile; <= 0: one GPU renders the frame as one tile, several GPUs of this
                                     process share ONE tile (device-side chunk stealing, ct_gpu_render_shared) */
    /* Optional cross-process tile stealing (one process per GPU): name of a POSIX shared-memory
     * counter shared by all ranks of the job, or NULL/"" for a process-local counter. */
    const char *shared_counter_name;
    int32_t rank, world_size;     /* only used with shared_counter_name */
    const char *gpu_library;      /* path of libct_gpu.so; NULL = next to libct_host.so */
} ct_host_boss_config;

typedef struct ct_host_frame_stats {
    ct_ray_counters rays;         /* this process's devices only */
    float device_ms_max;          /* max over this process's devices of the summed kernel time of the tiles it rendered (CUDA events) */
    double wall_ms;               /* dispatch -> bitmap complete */
    int32_t tiles_total, tiles_mine;
    uint64_t kernel_launches;     /* CUDA kernels launched for this frame by this process */
} ct_host_frame_stats;

/* RayThread's first-call half (raythread.cpp:647-654): builds the BVH if needed and uploads to every device. */
ct_host_boss *ct_host_boss_create(ct_host_scene *s, const ct_host_boss_config *cfg);
/* Launch device k's work on the caller's CUDA stream (cudaStream_t; e.g. the framework's current stream, so
 * that its events and collectives order with the tiles).  NULL = the library's own stream. */
int ct_host_boss_set_stream(ct_host_boss *b, int device_slot, void *cuda_stream);
/* HandleUpdates' camera half (raythread.cpp:557-572). */
int ct_host_boss_set_camera(ct_host_boss *b, const double pos[3], float yaw, float pitch, float roll);
/* One frame: dispatch all row tiles (dynamic stealing), wait, gather to devices[0], copy into bitmap
 * (may be NULL to skip the host copy).  With a shared counter only this rank's tiles are rendered and
 * copied; see ct_host_boss_tiles for who rendered what. */
int ct_host_boss_render(ct_host_boss *b, uint32_t *bitmap, int stride_pixels, ct_host_frame_stats *stats);
/* With a shared counter: exactly one rank calls this between frames, before the job-wide barrier that
 * precedes ct_host_boss_render (the counter must read 0 when the ranks start stealing). */
int ct_host_boss_reset_shared_counter(ct_host_boss *b);
/* Tiles rendered by this process in the last frame: writes up to max (y_start,y_end) pairs, returns count. */
int ct_host_boss_tiles(const ct_host_boss *b, int32_t *y_ranges, int max_tiles);
void ct_host_boss_destroy(ct_host_boss *b);

/* ---- controls + viewer tick (the interactive shell, minus the window) ----------------------- */
/* eventType_t (eventQueue.h:7-12) */
enum { CT_EVENT_MOUSE_MOVE = 0, CT_EVENT_MOUSE_CLICK = 1, CT_EVENT_KEY_UP = 2, CT_EVENT_KEY_DOWN = 3 };
/* eventManager_t (eventQueue.h:32-38) + the statics of HandleUpdates (raythread.cpp:548-552: changesMade, yaw,
 * pitch, roll) + scene->camera.position.  Host state only; needs no GPU. */
typedef struct ct_host_controls ct_host_controls;
/* Camera position from the scene file, yaw = pitch = roll = 0, "a frame is due".  event_capacity 0 = 1000. */
ct_host_controls *ct_host_controls_create(const ct_host_scene *s, uint32_t event_capacity);
/* AddEvent (eventQueue.cpp:5, called from the SDL loop cobbletrace.cpp:104): value = the key's character.
 * CT_ERR_INVALID when `capacity` events are pending (the reference overwrites unread slots instead). */
int ct_host_controls_add_event(ct_host_controls *c, uint32_t type, uint32_t value);
/* The decision half of HandleUpdates (raythread.cpp:557-562): `changesMade || HandleKeyboard(...)` -- the queue is
 * drained through HandleKeyboard (:388-434; w/s, d/a, i/o move by 0.1, y/p/r turn by pi/16, c = origin, m = log)
 * unless a frame is already due (so keys queued before the first frame are read on the second call).
 * Returns 1 when a frame has to be rendered now (and clears "due"), 0 when nothing changed, <0 on error. */
int ct_host_controls_update(ct_host_controls *c);
uint32_t ct_host_controls_pending(const ct_host_controls *c);     /* unread events */
uint64_t ct_host_controls_frames(const ct_host_controls *c);      /* how often update() returned 1 */
/* Current camera: position, (yaw, pitch, roll), rotation matrix (raythread.cpp:564-572).  Any pointer may be NULL. */
void ct_host_controls_camera(const ct_host_controls *c, double pos[3], float ypr[3], double rot[9]);
void ct_host_controls_destroy(ct_host_controls *c);

/* Blit's seam (draw2d.h:22-64): called once per tick with the bitmap, whether or not it changed. */
typedef void (*ct_host_present_fn)(void *user, const uint32_t *bitmap, int stride_pixels, int frame_is_new);
/* One iteration of the main loop (cobbletrace.cpp:88-118) minus SDL: RayThread -> HandleUpdates; when a frame is due
 * the camera goes to every device of the boss, the frame is rendered and read back into `bitmap` (blocking -- the
 * reference re-dispatches its workers and lets the bitmap fill in over the next ticks; here the tick ends with the
 * finished frame, i.e. what the reference's bitmap converges to); then `present` (may be NULL) gets the bitmap.
 * *rendered (may be NULL) = 1 when a new frame was rendered. */
int ct_host_viewer_tick(ct_host_boss *b, ct_host_controls *c, uint32_t *bitmap, int stride_pixels, ct_host_present_fn present,
                        void *user, int *rendered, ct_host_frame_stats *stats);

/* A ready-made frame sink for hosts without a window: writes the bitmap (0x00BBGGRR pixels, PutPixel draw2d.h:8-20) as a
 * binary PPM (P6).  Returns CT_OK or CT_ERR_INVALID with the reason in ct_host_last_error(). */
int ct_host_write_ppm(const char *path, const uint32_t *bitmap, int width, int height, int stride_pixels);

/* The tile dispenser on its own (what the boss steals from): a process-local atomic (shared_name NULL/"")
 * or a POSIX shared-memory counter common to all processes that open the same name. */
ct_host_tile_counter *ct_host_tile_counter_open(const char *shared_name);
int32_t ct_host_tile_counter_next(ct_host_tile_counter *c);       /* 0,1,2,... each value handed out exactly once */
void ct_host_tile_counter_reset(ct_host_tile_counter *c);
void ct_host_tile_counter_close(ct_host_tile_counter *c, int unlink_shared);

#ifdef __cplusplus
}
#endif
#endif /* CT_HOST_H */
