// oracle/ref_gpu_driver.cpp -- the reference with its workers replaced by libct_gpu.so: the INTEGRATION.md patch,
// compiled against the reference's own sources and run (test infrastructure; `make -C oracle ref_gpu`).
//
// This TU #includes a build-time generated copy of the reference's raythread.cpp (oracle/_ref/gen/raythread_gpu_gen.cpp)
// that differs from the one ct_ref uses (see ref_driver.cpp) in two more statements, both inside RayThread
// (raythread.cpp:641-666): AllocatePartitions -> CtGpuAllocatePartitions and a CtGpuRunPartitions call after
// HandleUpdates (oracle/ref_gpu_patch.h).  main() below does what cobbletrace.cpp's main does minus SDL: InitSceneData,
// ParseSceneFile, a calloc'd bitmap, then RayThread(&env, &scene) once per frame (cobbletrace.cpp:26-118).  Scene
// parsing, GetSceneTriangles, InitializeBVHState / BuildBVH, HandleKeyboard and HandleUpdates are the reference's
// code; the arrays they produce (bvh_node_t[], triangle_t[] with its 96-byte stride, light_t[]) go to
// ct_gpu_upload_scene as they are.
//
//   ct_ref_gpu --scene file.json [--chdir dir] [--width W --height H] [--depth D] [--threads N]
//              [--force-reflection R] [--keys STR] [--flags F] --frame OUT
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

int ct_sdl_stub_quiet = 1;
int g_ctMaxDepth = 10;
uint32_t g_ctGpuFlags = 0;

#define CT_HOOK_TRACERAY
#define CT_HOOK_CLOSEST
#define CT_HOOK_PIXEL(x, y)
#define CT_RAND() rand()

#include "raythread_gpu_gen.cpp"   // generated from /root/reference/raythread.cpp, includes ref_gpu_patch.h

int main(int argc, char **argv) {
    const char *sceneFile = NULL, *dir = NULL, *frameOut = NULL, *keys = NULL;
    int W = 640, H = 640, threads = 8;
    float forceReflection = -1;
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
        else if (ARG("--keys")) keys = argv[++i];
        else if (ARG("--flags")) g_ctGpuFlags = (uint32_t)atoi(argv[++i]);
        else { fprintf(stderr, "unknown arg %s\n", argv[i]); return 2; }
    }
    if (!sceneFile || !frameOut) { fprintf(stderr, "usage: ct_ref_gpu --scene file.json --frame out.bin [...]\n"); return 2; }
    char cwd[4096]; if (!getcwd(cwd, sizeof cwd)) return 2;
    char *frameP = (char *)malloc(strlen(cwd) + strlen(frameOut) + 2);
    if (frameOut[0] == '/') strcpy(frameP, frameOut); else sprintf(frameP, "%s/%s", cwd, frameOut);
    if (dir && chdir(dir) != 0) { perror("chdir"); return 2; }

    static scene_t scene;
    InitSceneData(&scene);
    ParseSceneFile((char *)sceneFile, &scene);
    scene.settings.supersampling = false;        // parity settings, as in ref_driver.cpp
    scene.settings.subsampling = false;
    scene.settings.numberOfThreads = threads;    // = partitions: each becomes one ct_gpu_render_tile call
    if (forceReflection >= 0)
        for (int i = 0; i < scene.objectStack.index; i++) scene.objectStack.objects[i].material.reflection = forceReflection;

    environment_t env = {};
    bitmapSettings_t bitmap = {};
    bitmap.memory = calloc((size_t)W * H, sizeof(uint32_t));      // cobbletrace.cpp:57
    bitmap.width = W;
    bitmap.height = H;
    env.bitmap = &bitmap;
    env.events.capacity = 1000;
    env.events.queue = (event_t *)calloc(env.events.capacity, sizeof(event_t));

    RayThread(&env, &scene);                     // first call: BVH build, upload, first frame (cobbletrace.cpp:114)
    if (keys) {
        // key presses only take effect from the second frame on (static changesMade = true, raythread.cpp:548,557)
        for (const char *k = keys; *k; k++) AddEvent(&env.events, {ET_KEY_DOWN, EM_NONE, {0, 0}, (uint32_t)*k});
        RayThread(&env, &scene);
    }
    FILE *f = fopen(frameP, "wb");
    if (!f || fwrite(bitmap.memory, 4, (size_t)W * H, f) != (size_t)W * H) { fprintf(stderr, "cannot write %s\n", frameP); return 4; }
    fclose(f);
    ct_gpu_shutdown(0);
    printf("{\"scene\": \"%s\", \"width\": %d, \"height\": %d, \"depth\": %d, \"partitions\": %d, \"triangle_stride\": %d}\n",
           sceneFile, W, H, g_ctMaxDepth, threads, (int)sizeof(triangle_t));
    return 0;
}
