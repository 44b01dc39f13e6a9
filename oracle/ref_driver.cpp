// oracle/ref_driver.cpp -- headless driver around the UNMODIFIED CobbleTrace
// reference sources (test infrastructure; never part of the product path).
//
// The reference keeps its whole hot path `static` inside raythread.cpp
// (TraceRay :353, ClosestIntersection :197, HandleUpdates :546, FINF :58), so
// this TU #includes a build-time generated copy of it (oracle/_ref/gen/
// raythread_gen.cpp, produced by oracle/Makefile with sed from
// /root/reference/raythread.cpp).  The generated copy differs from the
// original in exactly three, behaviour-neutral ways:
//   1. :119  `if (t < 0)` -> `if (*t < 0)`  (pointer-vs-int compare in the dead,
//      never-called IntersectRayTriangle; a hard error in g++ 13),
//   2. :480/:508 the recursion-depth literal `10` -> `g_ctMaxDepth` (harness
//      parameter; 10 by default, BASELINE config 3 uses 2),
//   3. empty-by-default CT_HOOK_* macros at the entry of TraceRay and
//      ClosestIntersection (ray-kind counters, only with -DCT_COUNT).
// bvh.cpp gets the same kind of hooks at IntersectAABB / IntersectTriangle.
//
// What it does: replays RayThread's first-call initialisation
// (raythread.cpp:647-654), then drives the reference's own boss/worker
// (AllocatePartitions :596, HandleUpdates :546, status polling :657-661).
//
// Modes (combine freely):
//   --frame F      write the W*H uint32 framebuffer (0x00BBGGRR) after one frame
//   --hits F       write per-pixel primary-ray hit records {u32 found,u32 idx,f32 t}
//   --dump-scene F write the flattened scene + BVH (format: see tests/ctscene.py)
//   --time K       time K frames through the boss/worker, print JSON
//   --counters     print ray / box-test / triangle-test counters (needs -DCT_COUNT build)
//   --supersampling-hash  settings.supersampling (:460-505: 4x4 jittered samples per pixel, running blend) with the jitter taken
//                  from a counter-based generator instead of rand() (see CT_RAND below)
//   --subsampling  leave settings.subsampling on (every other row traced, the rows between averaged, :512-531)
//   --keys STR     feed STR as key presses to the reference's HandleKeyboard (raythread.cpp:388) before
//                  the first frame: y/p/r rotate by pi/16, wasd/io move the camera by 0.1
//   --session STR  an interactive session: STR = batches of key presses separated by '|'; every batch is queued and one
//                  frame rendered (one main-loop tick of cobbletrace.cpp:88-118 with the workers run to completion);
//                  the camera and the frame's FNV-1a hash after every tick are printed as "session": [...]
//                  (tick 0 = the first frame, rendered before any key is read, raythread.cpp:548,557)

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <time.h>
#include <string>

int ct_sdl_stub_quiet = 1;
int g_ctMaxDepth = 10;

#ifdef CT_COUNT
unsigned long long g_ctTraceRayCalls = 0, g_ctClosestCalls = 0;
extern unsigned long long g_ctBoxTests, g_ctTriTests;
#define CT_HOOK_TRACERAY __atomic_fetch_add(&g_ctTraceRayCalls, 1ull, __ATOMIC_RELAXED);
#define CT_HOOK_CLOSEST  __atomic_fetch_add(&g_ctClosestCalls, 1ull, __ATOMIC_RELAXED);
#else
#define CT_HOOK_TRACERAY
#define CT_HOOK_CLOSEST
#endif

// Supersampling (raythread.cpp:460-505) draws its jitter from libc rand() on every worker thread at once, so the
// reference's own output is not reproducible.  The generated copy routes those two calls through
// CT_RAND(): rand() itself by default; with --supersampling-hash a counter-based generator keyed on (x, y, call
// number within the pixel) -- the only change, and the same function the oracle restatement and the GPU use.
int g_ctHashRand = 0;
static thread_local uint32_t t_ctPx, t_ctPy, t_ctPk;
static inline uint32_t CtHash3(uint32_t x, uint32_t y, uint32_t k) {
    uint32_t h = x * 0x9E3779B1u ^ (y * 0x85EBCA77u) ^ (k * 0xC2B2AE3Du);
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    return h;
}
#define CT_HOOK_PIXEL(x, y) t_ctPx = (uint32_t)(x); t_ctPy = (uint32_t)(y); t_ctPk = 0;
#define CT_RAND() (g_ctHashRand ? (int)(CtHash3(t_ctPx, t_ctPy, t_ctPk++) & 0x7FFFFFFFu) : rand())

#include "raythread_gen.cpp"   // generated from /root/reference/raythread.cpp

static double NowMs() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static bool AllFinished(scene_t *scene) {
    for (int i = 0; i < scene->settings.numberOfThreads; i++) {
        if (*(volatile worker_status_t *)&displayPart[i]->status != WS_FINISHED) return false;
    }
    return true;
}

// One frame through the reference's boss/worker; returns dispatch->all-finished ms.
static double RenderFrame(environment_t *env, scene_t *scene, bvh_state_t *bvh, bool first) {
    if (!first) {
        // HandleUpdates re-dispatches only when HandleKeyboard reports a change
        // (raythread.cpp:557-560); 'm' logs the camera and changes nothing (:424-429).
        AddEvent(&env->events, {ET_KEY_DOWN, EM_NONE, {0, 0}, (uint32_t)'m'});
    }
    double t0 = NowMs();
    HandleUpdates(env, scene, bvh);
    while (!AllFinished(scene)) { /* tight poll */ }
    return NowMs() - t0;
}

static void WriteFile(const char *path, const void *data, size_t bytes) {
    FILE *f = fopen(path, "wb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(4); }
    if (fwrite(data, 1, bytes, f) != bytes) { fprintf(stderr, "short write %s\n", path); exit(4); }
    fclose(f);
}

static void DumpScene(const char *path, scene_t *scene, bvh_state_t *bvh) {
    FILE *f = fopen(path, "wb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(4); }
    uint32_t nTri = bvh->triangles.size, nLights = scene->lightStack.index, nNodes = bvh->nodesUsed, zero = 0;
    fwrite("CTSCENE1", 1, 8, f);
    fwrite(&nTri, 4, 1, f); fwrite(&nLights, 4, 1, f); fwrite(&nNodes, 4, 1, f); fwrite(&zero, 4, 1, f);
    fwrite(&scene->camera.position, sizeof(double), 3, f);
    fwrite(&scene->camera.rotation.data[0][0], sizeof(double), 9, f);
    for (uint32_t k = 0; k < nTri; k++) {
        triangle_t *t = &bvh->triangles.data[k];
        fwrite(&t->p1, sizeof(double), 3, f); fwrite(&t->p2, sizeof(double), 3, f); fwrite(&t->p3, sizeof(double), 3, f);
    }
    for (uint32_t k = 0; k < nTri; k++) {
        material_t *m = &scene->objectStack.objects[scene->triangleLookup.indexes[k]].material;
        fwrite(&m->color, 4, 1, f); fwrite(&m->specular, 4, 1, f); fwrite(&m->reflection, 4, 1, f);
    }
    for (uint32_t k = 0; k < nLights; k++) {
        light_t *l = &scene->lightStack.lights[k];
        int32_t type = (int32_t)l->type;
        fwrite(&type, 4, 1, f); fwrite(&l->intensity, 4, 1, f);
        fwrite(&l->position, sizeof(double), 3, f); fwrite(&l->direction, sizeof(double), 3, f);
    }
    for (uint32_t k = 0; k < nNodes; k++) {
        bvh_node_t *n = &bvh->bvhNodes[k];
        fwrite(&n->aabbMin, sizeof(double), 3, f); fwrite(&n->aabbMax, sizeof(double), 3, f);
        fwrite(&n->leftNode, 4, 1, f); fwrite(&n->firstTriangleIndex, 4, 1, f); fwrite(&n->triangleCount, 4, 1, f);
        fwrite(&zero, 4, 1, f);
    }
    fwrite(bvh->triangles.indexes, 4, nTri, f);
    fclose(f);
}

struct hit_record_t { uint32_t found; uint32_t index; float t; };

// --kat IN OUT: known-answer vectors straight from the reference's two primitives (bvh.cpp:147,165).
// IN : u32 n, then org[n][3] dir[n][3] tri[n][9] bmin[n][3] bmax[n][3] (double), t[n] (float)
// OUT: tri_hit[n] box_hit[n] (u32), t_out[n] (float)
extern bool IntersectTriangle(ray_t *ray, triangle_t *triangle);
extern bool IntersectAABB(ray_t *ray, v3_t bmin, v3_t bmax);
static int RunKat(const char *in, const char *out) {
    FILE *f = fopen(in, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", in); return 4; }
    uint32_t n = 0;
    if (fread(&n, 4, 1, f) != 1) return 4;
    double *org = (double *)malloc(24ull * n), *dir = (double *)malloc(24ull * n), *tri = (double *)malloc(72ull * n);
    double *mn = (double *)malloc(24ull * n), *mx = (double *)malloc(24ull * n);
    float *t = (float *)malloc(4ull * n);
    if (fread(org, 24, n, f) != n || fread(dir, 24, n, f) != n || fread(tri, 72, n, f) != n || fread(mn, 24, n, f) != n ||
        fread(mx, 24, n, f) != n || fread(t, 4, n, f) != n) { fprintf(stderr, "short KAT input\n"); return 4; }
    fclose(f);
    uint32_t *th = (uint32_t *)malloc(4ull * n), *bh = (uint32_t *)malloc(4ull * n);
    for (uint32_t i = 0; i < n; i++) {
        ray_t ray = {{org[3 * i], org[3 * i + 1], org[3 * i + 2]}, {dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]}, t[i]};
        v3_t bmin = {mn[3 * i], mn[3 * i + 1], mn[3 * i + 2]}, bmax = {mx[3 * i], mx[3 * i + 1], mx[3 * i + 2]};
        bh[i] = IntersectAABB(&ray, bmin, bmax) ? 1u : 0u;
        triangle_t tr = {};
        tr.p1 = {tri[9 * i], tri[9 * i + 1], tri[9 * i + 2]};
        tr.p2 = {tri[9 * i + 3], tri[9 * i + 4], tri[9 * i + 5]};
        tr.p3 = {tri[9 * i + 6], tri[9 * i + 7], tri[9 * i + 8]};
        th[i] = IntersectTriangle(&ray, &tr) ? 1u : 0u;
        t[i] = ray.t;
    }
    f = fopen(out, "wb");
    if (!f) return 4;
    fwrite(th, 4, n, f); fwrite(bh, 4, n, f); fwrite(t, 4, n, f);
    fclose(f);
    return 0;
}

int main(int argc, char **argv) {
    const char *sceneFile = NULL, *dir = NULL, *frameOut = NULL, *hitsOut = NULL, *sceneOut = NULL, *keys = NULL, *session = NULL;
    int W = 640, H = 640, threads = 8, timeFrames = 0, warmFrames = 0;
    float forceReflection = -1;
    bool counters = false, subsampling = false, supersampling = false;
    if (argc == 4 && strcmp(argv[1], "--kat") == 0) return RunKat(argv[2], argv[3]);
    for (int i = 1; i < argc; i++) {
        #define ARG(name) (strcmp(argv[i], name) == 0 && i + 1 < argc)
        if (ARG("--scene")) sceneFile = argv[++i];
        else if (ARG("--chdir")) dir = argv[++i];
        else if (ARG("--width")) W = atoi(argv[++i]);
        else if (ARG("--height")) H = atoi(argv[++i]);
        else if (ARG("--depth")) g_ctMaxDepth = atoi(argv[++i]);
        else if (ARG("--threads")) threads = atoi(argv[++i]);
        else if (ARG("--force-reflection")) forceReflection = (float)atof(argv[++i]);
        else if (ARG("--frame")) frameOut = argv[++i];
        else if (ARG("--hits")) hitsOut = argv[++i];
        else if (ARG("--dump-scene")) sceneOut = argv[++i];
        else if (ARG("--time")) timeFrames = atoi(argv[++i]);
        else if (ARG("--warmup")) warmFrames = atoi(argv[++i]);   // untimed frames before the --time frames
        else if (ARG("--session")) session = argv[++i];
        else if (ARG("--keys")) keys = argv[++i];       // key presses fed to HandleKeyboard before the first frame
        else if (strcmp(argv[i], "--counters") == 0) counters = true;
        else if (strcmp(argv[i], "--subsampling") == 0) subsampling = true;   // settings.subsampling (raythread.cpp:512-531); use --threads 1
        else if (strcmp(argv[i], "--supersampling-hash") == 0) { supersampling = true; g_ctHashRand = 1; }   // :460-505 with the counter-based CT_RAND
        else if (strcmp(argv[i], "--verbose") == 0) ct_sdl_stub_quiet = 0;
        else { fprintf(stderr, "unknown arg %s\n", argv[i]); return 2; }
    }
    if (!sceneFile) { fprintf(stderr, "usage: ct_ref --scene file.json [--chdir dir] ...\n"); return 2; }
    char *frameAbs = frameOut ? realpath(".", NULL) : NULL; (void)frameAbs;
    // Output paths are resolved before chdir so they may be relative to the caller's cwd.
    char cwd[4096]; if (!getcwd(cwd, sizeof cwd)) return 2;
    auto absPath = [&](const char *p) -> char * {
        if (!p) return NULL;
        if (p[0] == '/') return strdup(p);
        char *r = (char *)malloc(strlen(cwd) + strlen(p) + 2);
        sprintf(r, "%s/%s", cwd, p);
        return r;
    };
    char *frameP = absPath(frameOut), *hitsP = absPath(hitsOut), *sceneP = absPath(sceneOut);
    if (dir && chdir(dir) != 0) { perror("chdir"); return 2; }

    double tLoad0 = NowMs();
    static scene_t scene;
    InitSceneData(&scene);
    ParseSceneFile((char *)sceneFile, &scene);
    double tLoad1 = NowMs();

    // Parity settings: sampling modes off (supersampling draws from shared rand()).
    scene.settings.supersampling = supersampling;   // only with --supersampling-hash (deterministic jitter)
    scene.settings.subsampling = subsampling;   // deterministic with one thread (several threads race on the rows between partitions)
    scene.settings.numberOfThreads = threads;
    if (forceReflection >= 0) {
        for (int i = 0; i < scene.objectStack.index; i++) scene.objectStack.objects[i].material.reflection = forceReflection;
    }

    // RayThread's first-call sequence, raythread.cpp:647-654.
    uint32_t size;
    triangle_t *triangles = GetSceneTriangles(&scene, &size);
    double tBuild0 = NowMs();
    static bvh_state_t bvh;
    bvh = InitializeBVHState(triangles, size);
    BuildBVH(&bvh);
    double tBuild1 = NowMs();

    environment_t env = {};
    bitmapSettings_t bitmap = {};
    bitmap.memory = calloc((size_t)W * H, sizeof(uint32_t));
    bitmap.width = W;
    bitmap.height = H;
    env.bitmap = &bitmap;
    env.events.capacity = 1000;
    env.events.queue = (event_t *)calloc(env.events.capacity, sizeof(event_t));

    AllocatePartitions(&scene);
    double ms = RenderFrame(&env, &scene, &bvh, true);
    if (keys) {
        // HandleKeyboard is short-circuited on the very first HandleUpdates (static changesMade = true,
        // raythread.cpp:548,557), so key presses only take effect from the second frame on.
        for (const char *k = keys; *k; k++) AddEvent(&env.events, {ET_KEY_DOWN, EM_NONE, {0, 0}, (uint32_t)*k});
        ms = RenderFrame(&env, &scene, &bvh, false);
    }

    std::string sessionLog;
    if (session) {
        const char *k = session;
        for (int tick = 0;; tick++) {
            if (tick > 0) {
                for (; *k && *k != '|'; k++) AddEvent(&env.events, {ET_KEY_DOWN, EM_NONE, {0, 0}, (uint32_t)*k});
                ms = RenderFrame(&env, &scene, &bvh, false);
            }
            uint64_t h = 1469598103934665603ull;
            const uint32_t *px = (const uint32_t *)bitmap.memory;
            for (size_t i = 0; i < (size_t)W * H; i++) h = (h ^ px[i]) * 1099511628211ull;
            char buf[1024];
            const double *r = &scene.camera.rotation.data[0][0];
            snprintf(buf, sizeof buf, "%s{\"pos\": [%.17g, %.17g, %.17g], \"rot\": [%.17g, %.17g, %.17g, %.17g, %.17g, %.17g, %.17g, %.17g, %.17g], \"fnv\": \"%016llx\"}",
                     tick ? ", " : "", scene.camera.position.x, scene.camera.position.y, scene.camera.position.z,
                     r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8], (unsigned long long)h);
            sessionLog += buf;
            if (tick > 0) { if (*k == '|') k++; else break; }
            else if (!*k) break;
        }
    }

    if (frameP) WriteFile(frameP, bitmap.memory, (size_t)W * H * 4);
    if (sceneP) DumpScene(sceneP, &scene, &bvh);

#ifdef CT_COUNT
    unsigned long long frameTrace = g_ctTraceRayCalls, frameClosest = g_ctClosestCalls, frameBox = g_ctBoxTests, frameTri = g_ctTriTests;
#endif

    if (hitsP) {
        // Primary-ray hit records, same loops/mapping as RayTracePartition (:454-457, :507)
        // and the two lines of ClosestIntersection that produce the index (:204-207).
        hit_record_t *hits = (hit_record_t *)calloc((size_t)W * H, sizeof(hit_record_t));
        for (size_t i = 0; i < (size_t)W * H; i++) hits[i].found = 0xFFFFFFFFu; // never traced / dropped
        viewport_t vp = {1, 1, 1};
        float width = bitmap.height / 2;
        for (int part = 0; part < scene.settings.numberOfThreads; part++)   // exactly the rows the workers traced
        for (int x = -width; x < width; x++) {
            for (int y = displayPart[part]->yStart; y < displayPart[part]->yEnd; y++) {
                v3_t direction = CanvasToViewport(&bitmap, vp, {(float)x, (float)y}) * scene.camera.rotation;
                ray_t ray = {scene.camera.position, direction, 1e30f};
                float tclosest = FINF;
                uint32_t closestIndex = 0;
                IntersectBVHClosest(&ray, bvh.rootNodeIdx, &bvh, &tclosest, &closestIndex);
                int col = x + bitmap.width / 2, row = bitmap.height / 2 - y;
                if (row < 0 || row >= bitmap.height || col < 0 || col >= bitmap.width) continue;
                hit_record_t *h = &hits[(size_t)row * W + col];
                h->found = (ray.t != 1e30f) ? 1u : 0u;
                h->index = closestIndex;
                h->t = tclosest;
            }
        }
        WriteFile(hitsP, hits, (size_t)W * H * sizeof(hit_record_t));
        free(hits);
    }

    double best = ms, sum = 0;
    if (timeFrames > 0) {
        for (int k = 0; k < warmFrames; k++) RenderFrame(&env, &scene, &bvh, false);
        for (int k = 0; k < timeFrames; k++) {
            double m = RenderFrame(&env, &scene, &bvh, false);
            sum += m;
            if (k == 0 || m < best) best = m;
        }
    }

    printf("{\"scene\": \"%s\", \"width\": %d, \"height\": %d, \"depth\": %d, \"threads\": %d, "
           "\"triangles\": %u, \"lights\": %d, \"nodes\": %u, \"load_ms\": %.3f, \"build_ms\": %.3f, "
           "\"first_frame_ms\": %.3f, \"frames_timed\": %d, \"best_ms\": %.3f, \"mean_ms\": %.3f",
           sceneFile, W, H, g_ctMaxDepth, threads, size, scene.lightStack.index, bvh.nodesUsed,
           tLoad1 - tLoad0, tBuild1 - tBuild0, ms, timeFrames, best, timeFrames > 0 ? sum / timeFrames : ms);
#ifdef CT_COUNT
    if (counters) {
        // primary = TraceRay calls from the pixel loop (H*H), reflection = recursive TraceRay calls,
        // shadow = ClosestIntersection calls made by ComputeLighting (:304).
        unsigned long long primary = (unsigned long long)(2 * (H / 2)) * (unsigned long long)(2 * (H / 2));
        // with numberOfThreads not dividing H, fewer rows are traced (:576); count from partitions instead
        unsigned long long rows = 0;
        for (int i = 0; i < scene.settings.numberOfThreads; i++) rows += displayPart[i]->yEnd - displayPart[i]->yStart;
        primary = rows * (unsigned long long)(2 * (H / 2));
        printf(", \"rays_primary\": %llu, \"rays_reflection\": %llu, \"rays_shadow\": %llu, \"box_tests\": %llu, \"tri_tests\": %llu",
               primary, frameTrace - primary, frameClosest - frameTrace, frameBox, frameTri);
    }
#else
    (void)counters;
#endif
    if (session) printf(", \"session\": [%s]", sessionLog.c_str());
    printf("}\n");
    fflush(stdout);
    _exit(0); // worker threads never return (raythread.cpp:438)
}
