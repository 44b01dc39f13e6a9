// ct_viewer.cpp -- the interactive shell around the tracer, without the window.
//
// One iteration of the reference's main loop (cobbletrace.cpp:88-118) is: turn pending SDL key presses into
// queue entries (AddEvent, :104), RayThread -> HandleUpdates (raythread.cpp:639-665, :546-590: drain the queue
// through HandleKeyboard :388-434, rebuild the camera matrix and re-dispatch the workers only when something
// changed), Blit the bitmap (draw2d.h:22-64), sleep 33 ms.  Here:
//   Controls     = eventManager_t (eventQueue.h:32-38) + HandleUpdates' statics (changesMade, yaw, pitch, roll)
//                  + scene->camera.position.  Pure host state, no GPU.
//   viewer_tick  = RayThread + Blit for one iteration: update the controls; if a frame is due, set the camera on
//                  every device, render and read back through the boss (blocking: a frame is milliseconds, the
//                  reference's tick is 33 ms); then hand the bitmap to the caller's `present` hook -- the seam
//                  where SDL / GL interop would go (there is no SDL in this image; the window itself is out of scope).
// Quirks kept, because they decide which frame is on screen:
//   * HandleUpdates' `changesMade || HandleKeyboard(...)` short-circuits on the very first call: keys queued
//     before the first frame are read on the SECOND tick (raythread.cpp:548,557);
//   * 'm' only logs the camera but still reports a change (:424-429) -> the frame is rendered again;
//   * yaw/pitch/roll are floats stepped by the double M_PI_4/4; positions are doubles stepped by 0.1;
//   * 'c' moves the camera to the origin (not to the scene file's position).
// One deliberate difference: the reference's ring buffer overwrites unread events once more than `capacity` are
// pending and then replays stale slots (eventQueue.cpp:5-12 never checks `active`); add_event refuses instead.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "ct_host.h"
#include "ct_scene.hpp"

namespace cth {

struct Boss;
void boss_set_camera(Boss *b, const double pos[3], float yaw, float pitch, float roll);
void boss_render(Boss *b, uint32_t *bitmap, int stride, ct_host_frame_stats *stats);

struct QueuedEvent {
    uint32_t type;
    uint32_t value;
};

struct Controls {
    std::vector<QueuedEvent> ring;
    uint32_t head = 0, pending = 0;          // oldest unread slot, number of unread events
    bool frame_due = true;                   // HandleUpdates' `static bool changesMade = true`
    float yaw = 0, pitch = 0, roll = 0;      // HandleUpdates' statics (:549-551)
    Vec3 pos;                                // scene->camera.position
    uint64_t frames = 0;                     // frames dispatched so far
};

Controls *controls_create(const Scene *scene, uint32_t capacity) {
    auto *c = new Controls();
    c->ring.resize(capacity ? capacity : 1000);     // EVENT_QUEUE_SIZE, cobbletrace.cpp
    c->pos = scene->cam_pos;
    return c;
}

bool controls_add_event(Controls *c, uint32_t type, uint32_t value) {
    if (c->pending == c->ring.size()) return false;
    c->ring[(c->head + c->pending) % c->ring.size()] = {type, value};
    c->pending++;
    return true;
}

// HandleKeyboard (raythread.cpp:388-434): every queued event is consumed; true when any of them is a bound key.
static bool drain_keys(Controls *c) {
    const double kStep = 0.1, kTurn = M_PI_4 / 4;
    bool changed = false;
    while (c->pending) {
        const QueuedEvent e = c->ring[c->head];
        c->head = (c->head + 1) % c->ring.size();
        c->pending--;
        if (e.type != CT_EVENT_KEY_DOWN) continue;
        bool bound = true;
        switch (e.value) {
            case 'w': c->pos.y += kStep; break;
            case 's': c->pos.y -= kStep; break;
            case 'd': c->pos.x += kStep; break;
            case 'a': c->pos.x -= kStep; break;
            case 'i': c->pos.z += kStep; break;
            case 'o': c->pos.z -= kStep; break;
            case 'r': c->roll = (float)(c->roll + kTurn); break;      // float += double
            case 'y': c->yaw = (float)(c->yaw + kTurn); break;
            case 'p': c->pitch = (float)(c->pitch + kTurn); break;
            case 'c': c->pos = {0, 0, 0}; c->yaw = c->pitch = c->roll = 0; break;
            case 'm': fprintf(stderr, "camera position: %f %f %f\n", c->pos.x, c->pos.y, c->pos.z); break;
            default: bound = false;
        }
        changed = changed || bound;
    }
    return changed;
}

// The decision half of HandleUpdates (:557-572).  Returns true when a frame has to be dispatched.
bool controls_update(Controls *c) {
    if (!c->frame_due) c->frame_due = drain_keys(c);     // `changesMade || HandleKeyboard(...)`: no drain while a frame is due
    if (!c->frame_due) return false;
    c->frame_due = false;
    c->frames++;
    return true;
}

void controls_camera(const Controls *c, double pos[3], float ypr[3], double rot[9]) {
    if (pos) { pos[0] = c->pos.x; pos[1] = c->pos.y; pos[2] = c->pos.z; }
    if (ypr) { ypr[0] = c->yaw; ypr[1] = c->pitch; ypr[2] = c->roll; }
    if (rot) camera_rotation(c->yaw, c->pitch, c->roll, rot);
}

uint32_t controls_pending(const Controls *c) { return c->pending; }
uint64_t controls_frames(const Controls *c) { return c->frames; }
void controls_destroy(Controls *c) { delete c; }

bool viewer_tick(Boss *b, Controls *c, uint32_t *bitmap, int stride, ct_host_present_fn present, void *user, ct_host_frame_stats *stats) {
    const bool render = controls_update(c);
    if (render) {
        double pos[3];
        controls_camera(c, pos, nullptr, nullptr);
        boss_set_camera(b, pos, c->yaw, c->pitch, c->roll);
        boss_render(b, bitmap, stride, stats);
    }
    if (present) present(user, bitmap, stride, render ? 1 : 0);      // Blit happens every tick, new frame or not
    return render;
}

}  // namespace cth
