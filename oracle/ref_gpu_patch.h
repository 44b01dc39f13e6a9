/* oracle/ref_gpu_patch.h -- the reference-side binding of include/ct_gpu.h, as a maintainer of jeng/CobbleTrace would
 * add it to raythread.cpp (this is the code INTEGRATION.md shows; here it is compiled and run).
 *
 * Test infrastructure: oracle/Makefile (`make ref_gpu`) generates a copy of the reference's raythread.cpp in which
 * exactly two statements of RayThread (raythread.cpp:641-666) are replaced,
 *
 *     AllocatePartitions(scene);                 ->   CtGpuAllocatePartitions(env, scene, &bvhState);     (:653)
 *     HandleUpdates(env, scene, &bvhState);      ->   HandleUpdates(...); CtGpuRunPartitions(env, scene); (:664)
 *
 * and #includes this header in front of RayThread.  Everything else is the reference's own code: ParseSceneFile,
 * GetSceneTriangles, InitializeBVHState, BuildBVH, HandleKeyboard and HandleUpdates (camera matrix :564-572, partition
 * arithmetic :574-588).  The worker threads (RayTracePartition :437-543) are what the GPU replaces: partitions are
 * allocated without SDL threads, and after HandleUpdates has marked them WS_READY each [yStart, yEnd) goes through
 * ct_gpu_render_tile; ct_gpu_readback then fills bitmap->memory, where the workers' PutPixel would have written.
 * The reference's arrays are passed as they are: bvh_node_t[], triangle_t[] at its 96-byte stride, light_t[]. */
#ifndef CT_REF_GPU_PATCH_H
#define CT_REF_GPU_PATCH_H

#include "ct_gpu.h"

static_assert(sizeof(bvh_node_t) == sizeof(ct_bvh_node), "bvh_node_t layout (bvh.h:5-11)");
static_assert(sizeof(material_t) == sizeof(ct_material), "material_t layout (scenefile.h:25-29)");
static_assert(sizeof(light_t) == sizeof(ct_light), "light_t layout (scenefile.h:61-66)");
static_assert(sizeof(triangle_t) == 96, "triangle_t stride (scenefile.h:36-41)");

extern int g_ctMaxDepth;        /* the harness's recursion depth; the reference's literal is 10 (raythread.cpp:508) */
extern uint32_t g_ctGpuFlags;   /* CT_FLAG_* for the upload (0 by default) */

static void CtGpuFail(const char *what) {
    fprintf(stderr, "%s: %s\n", what, ct_gpu_last_error());
    exit(4);
}

/* Stands where AllocatePartitions(scene) stood (raythread.cpp:653): the same partition records, no worker threads,
 * and the scene goes to the device. */
static void CtGpuAllocatePartitions(environment_t *env, scene_t *scene, bvh_state_t *bvhState) {
    displayPart = (display_partition_t **)calloc(scene->settings.numberOfThreads, sizeof(display_partition_t *));
    for (int i = 0; i < scene->settings.numberOfThreads; i++) {
        displayPart[i] = (display_partition_t *)calloc(1, sizeof(display_partition_t));
        if (displayPart[i] == NULL) exit(2);
        displayPart[i]->status = WS_FINISHED;
        displayPart[i]->thread = (SDL_Thread *)displayPart[i];      /* non-NULL: HandleUpdates exits on NULL (:590) */
    }
    /* per-triangle material, in GetSceneTriangles order: objects[triangleLookup.indexes[k]].material (:211) */
    uint32_t n = bvhState->triangles.size;
    ct_material *mats = (ct_material *)malloc((size_t)n * sizeof *mats);
    for (uint32_t k = 0; k < n; k++)
        memcpy(&mats[k], &scene->objectStack.objects[scene->triangleLookup.indexes[k]].material, sizeof *mats);
    ct_scene_desc d;
    memset(&d, 0, sizeof d);
    d.struct_size = sizeof d;
    d.flags = g_ctGpuFlags;
    d.n_triangles = n;
    d.triangle_stride = sizeof(triangle_t);                         /* p1,p2,p3 at offset 0, centroid skipped */
    d.triangles = bvhState->triangles.data;
    d.materials = mats;
    d.n_nodes = bvhState->nodesUsed;
    d.nodes = (const ct_bvh_node *)bvhState->bvhNodes;
    d.tri_indexes = bvhState->triangles.indexes;
    d.n_lights = (uint32_t)scene->lightStack.index;
    d.lights = (const ct_light *)scene->lightStack.lights;
    memcpy(d.camera_position, &scene->camera.position, sizeof d.camera_position);
    memcpy(d.camera_rotation, scene->camera.rotation.data, sizeof d.camera_rotation);
    d.viewport[0] = d.viewport[1] = d.viewport[2] = 1.0f;            /* :554 */
    d.width = env->bitmap->width; d.height = env->bitmap->height;
    d.max_depth = g_ctMaxDepth;
    d.background = BACKGROUND_COLOR;                                 /* :59 */
    if (ct_gpu_upload_scene(0, &d) != CT_OK) CtGpuFail("ct_gpu_upload_scene");
    free(mats);                                                      /* the library copied everything */
}

/* Stands after HandleUpdates (raythread.cpp:664), in the workers' place: every partition HandleUpdates marked
 * WS_READY is rendered with the camera it just computed, then the bitmap is read back. */
static void CtGpuRunPartitions(environment_t *env, scene_t *scene) {
    bool any = false;
    for (int i = 0; i < scene->settings.numberOfThreads; i++) {
        if (displayPart[i]->status != WS_READY) continue;
        if (!any && ct_gpu_set_camera(0, (const double *)&scene->camera.position, (const double *)scene->camera.rotation.data) != CT_OK)
            CtGpuFail("ct_gpu_set_camera");
        any = true;
        if (ct_gpu_render_tile(0, displayPart[i]->yStart, displayPart[i]->yEnd, NULL) != CT_OK) CtGpuFail("ct_gpu_render_tile");
        displayPart[i]->status = WS_FINISHED;
    }
    if (any && ct_gpu_readback(0, (uint32_t *)env->bitmap->memory, env->bitmap->width, 0, env->bitmap->height) != CT_OK)
        CtGpuFail("ct_gpu_readback");
}

#endif
