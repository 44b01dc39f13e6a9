/* examples/headless_viewer.c -- the reference's main loop (cobbletrace.cpp:88-118) against this library, in plain C:
 * load a scene file, build the BVH, drive one GPU through the boss, feed key presses, write every new frame as a PPM.
 *
 *   gcc -std=c99 examples/headless_viewer.c -Iinclude -Lcobbletrace_b200 -lct_host -Wl,-rpath,$PWD/cobbletrace_b200 -o viewer
 *   ./viewer scene_file_cube.json "yp|wd|oooo" out_prefix
 *
 * Key batches are separated by '|', one main-loop tick each (an empty batch renders nothing).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ct_host.h"

#define WIDTH 640   /* cobbletrace.cpp:14-15 */
#define HEIGHT 640

struct sink { const char *prefix; int frames; };

static void present(void *user, const uint32_t *bitmap, int stride_pixels, int frame_is_new) {   /* where Blit stands */
    struct sink *s = (struct sink *)user;
    char path[512];
    if (!frame_is_new) return;
    snprintf(path, sizeof path, "%s_%03d.ppm", s->prefix, s->frames++);
    if (ct_host_write_ppm(path, bitmap, WIDTH, HEIGHT, stride_pixels) != CT_OK) fprintf(stderr, "%s\n", ct_host_last_error());
}

int main(int argc, char **argv) {
    const char *keys = argc > 2 ? argv[2] : "";
    struct sink out = {argc > 3 ? argv[3] : "frame", 0};
    ct_host_scene *scene;
    ct_host_boss_config cfg;
    ct_host_boss *boss;
    ct_host_controls *ctl;
    uint32_t *bitmap, flags = 0;
    int tick = 0;
    if (argc < 2) { fprintf(stderr, "usage: %s scene.json [keys] [out_prefix]\n", argv[0]); return 2; }
    scene = ct_host_scene_load(argv[1], NULL);
    if (!scene) { fprintf(stderr, "%s\n", ct_host_last_error()); return 1; }
    ct_host_scene_render_flags(scene, &flags);               /* honour the file's sub-/supersampling settings */
    memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = sizeof cfg;
    cfg.width = WIDTH; cfg.height = HEIGHT; cfg.max_depth = 10; cfg.flags = flags;
    cfg.n_devices = 1; cfg.devices[0] = 0;
    boss = ct_host_boss_create(scene, &cfg);                  /* builds the BVH, uploads: RayThread's first-call block */
    if (!boss) { fprintf(stderr, "%s\n", ct_host_last_error()); return 1; }
    ctl = ct_host_controls_create(scene, 0);
    bitmap = (uint32_t *)calloc((size_t)WIDTH * HEIGHT, sizeof *bitmap);
    for (int more = 1; more;) {                               /* one iteration = one tick of the reference's loop */
        int fresh = 0;
        if (ct_host_viewer_tick(boss, ctl, bitmap, WIDTH, present, &out, &fresh, NULL) != CT_OK) { fprintf(stderr, "%s\n", ct_host_last_error()); return 1; }
        printf("tick %d: %s\n", tick++, fresh ? "new frame" : "nothing changed");
        more = tick == 1 ? *keys != 0 : *keys == '|';          /* "a|b" = two batches; "a|" = a batch and an idle tick */
        if (tick > 1 && more) keys++;
        for (; *keys && *keys != '|'; keys++) ct_host_controls_add_event(ctl, CT_EVENT_KEY_DOWN, (uint32_t)*keys);
    }
    free(bitmap);
    ct_host_controls_destroy(ctl);
    ct_host_boss_destroy(boss);
    ct_host_scene_free(scene);
    return 0;
}
