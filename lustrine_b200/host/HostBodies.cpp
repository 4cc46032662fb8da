// HostBodies.cpp — the host-side rigid-body hook (see include/lustrine/BulletPhysics.hpp).
// Bodies are integrated kinematically on the host; what matters to the particle step is
//   (1) simulate_bullet() publishes player_position before the particle predict, and
//   (2) the <= 100 particles inside the player's AABB, lowest reference slot first, are found by
//       the device compaction lgpu_aabb_first_k instead of the reference's O(N) host scan
//       (src/BulletPhysics.cpp:602-652) and parked on the proxy bodies.
#include <cmath>
#include <iostream>

#include "lgpu.h"
#include "lustrine/BulletPhysics.hpp"
#include "lustrine/RigidBodyHooks.hpp"

namespace Lustrine {
namespace Bullet {

static int new_body(Simulation* s, glm::vec3 position, glm::vec3 half, bool dynamic, bool detector) {
    Body b;
    b.position = position; b.half_extents = half; b.dynamic = dynamic; b.detector = detector; b.gravity = s->gravity;
    s->bodies.push_back(b);
    s->bodies_collisions.emplace_back();
    return s->num_bodies++;
}

void init_bullet(Simulation* s) { s->bodies.clear(); s->bodies_collisions.clear(); s->num_bodies = 0; }
void clean_bullet(Simulation* s) { s->bodies.clear(); s->bodies_collisions.clear(); s->num_bodies = 0; s->allocated_particles_bounding_boxes = false; }
void set_gravity(Simulation* s, glm::vec3 g) { s->gravity = g; for (Body& b : s->bodies) b.gravity = g; }
glm::vec3 get_gravity(Simulation* s) { return s->gravity; }
int add_box(Simulation* s, glm::vec3 position, bool is_dynamic) { return new_body(s, position, glm::vec3(0.5f), is_dynamic, false); }
int add_box(Simulation* s, glm::vec3 position, bool is_dynamic, glm::vec3 half) { return new_body(s, position, half, is_dynamic, false); }
int add_capsule(Simulation* s, glm::vec3 position, float radius, float height) { return new_body(s, position, glm::vec3(radius, 0.5f * height + radius, radius), true, false); }
int add_detector_block(Simulation* s, glm::vec3 position, glm::vec3 half) { return new_body(s, position, half, false, true); }

void allocate_particles_colliders(Simulation* s, int num_particles, float radius) {  // src/BulletPhysics.cpp:311-342
    if (s->allocated_particles_bounding_boxes) return;
    s->ptr_bounding_box_start = s->num_bodies;
    for (int i = 0; i < num_particles; i++) new_body(s, glm::vec3(0.0f), glm::vec3(radius), false, false);
    s->ptr_bounding_box_end = s->num_bodies;
    s->allocated_particles_bounding_boxes = true;
}
void bind_foreign_sand_positions(Simulation* s, glm::vec3* positions) { s->foreign_sand_positions = positions; }
void enable_particles_bounding_boxes(Simulation* s) { s->bounding_box_activated = true; }
void disable_particles_bounding_boxes(Simulation* s) { s->bounding_box_activated = false; }

void set_particles_box_colliders_positions(Simulation* s, glm::vec3*, int, int) {
    if ((int)s->bodies.size() > s->player_id && s->player_id >= 0) s->player_position = s->bodies[s->player_id].position;
    if (!s->bounding_box_activated || !s->gpu || s->ptr_bounding_box_end <= s->ptr_bounding_box_start) return;
    // particle_collide_with_player, src/BulletPhysics.cpp:596-600: |p - player| <= (box_scale + 4 r) / 2 per axis
    const glm::vec3 scale = s->player_box_scale + glm::vec3(s->particle_radius * 4.0f);
    const float center[3] = {s->player_position.x, s->player_position.y, s->player_position.z};
    const float half[3] = {scale.x * 0.5f, scale.y * 0.5f, scale.z * 0.5f};
    const int k = (int)(s->ptr_bounding_box_end - s->ptr_bounding_box_start);
    std::vector<float> out(3 * (size_t)k);
    int n = 0;
    if (lgpu_aabb_first_k(static_cast<lgpu_ctx*>(s->gpu), center, half, k, out.data(), &n) != LGPU_OK) return;
    for (int i = 0; i < k; i++) {
        Body& b = s->bodies[s->ptr_bounding_box_start + i];
        b.collision_response = i < n;
        if (i < n) { b.position = glm::vec3(out[3 * i], out[3 * i + 1], out[3 * i + 2]); b.velocity = glm::vec3(0.0f); }
    }
}

void simulate_bullet(Simulation* s, float dt, int sand_start, int sand_end) {  // src/BulletPhysics.cpp:163-174
    if (s->particles_bounding_box_current_state != s->particles_bounding_box_requested_state) {
        s->bounding_box_activated = s->particles_bounding_box_requested_state;
        s->particles_bounding_box_current_state = s->particles_bounding_box_requested_state;
    }
    set_particles_box_colliders_positions(s, s->foreign_sand_positions, sand_start, sand_end);
    for (Body& b : s->bodies) {
        if (!b.dynamic) continue;
        b.velocity += b.gravity * dt;
        b.velocity *= 1.0f / (1.0f + dt * b.linear_damping);
        b.position += b.velocity * dt;
        if (b.position.y < b.half_extents.y) { b.position.y = b.half_extents.y; if (b.velocity.y < 0.0f) b.velocity.y = 0.0f; }
    }
}

static bool overlap(const Body& a, const Body& b) {
    return std::fabs(a.position.x - b.position.x) <= a.half_extents.x + b.half_extents.x &&
           std::fabs(a.position.y - b.position.y) <= a.half_extents.y + b.half_extents.y &&
           std::fabs(a.position.z - b.position.z) <= a.half_extents.z + b.half_extents.z;
}
bool check_collision(Simulation* s, int a, int b) {
    if (a < 0 || b < 0 || a >= s->num_bodies || b >= s->num_bodies) return false;
    return overlap(s->bodies[a], s->bodies[b]);
}
bool do_collide(Simulation* s, int body) {
    for (int i = 0; i < s->num_bodies; i++) if (i != body && check_collision(s, body, i)) return true;
    return false;
}
int get_num_bodies(Simulation* s) { return s->num_bodies; }
void apply_impulse(Simulation* s, int body, glm::vec3 impulse, glm::vec3) { if (body >= 0 && body < s->num_bodies) s->bodies[body].velocity += impulse; }
glm::vec3 get_body_position(Simulation* s, int body) { return (body >= 0 && body < s->num_bodies) ? s->bodies[body].position : glm::vec3(0.0f); }
glm::vec3 get_body_velocity(Simulation* s, int body) { return (body >= 0 && body < s->num_bodies) ? s->bodies[body].velocity : glm::vec3(0.0f); }
void set_body_position(Simulation* s, int body, glm::vec3 p) { if (body >= 0 && body < s->num_bodies) s->bodies[body].position = p; }
void set_body_velocity(Simulation* s, int body, glm::vec3 v) { if (body >= 0 && body < s->num_bodies) s->bodies[body].velocity = v; }
void add_body_velocity(Simulation* s, int body, glm::vec3 v) { if (body >= 0 && body < s->num_bodies) s->bodies[body].velocity += v; }
static Body* body_at(Simulation* s, int id) { return (id >= 0 && id < s->num_bodies) ? &s->bodies[id] : nullptr; }
void set_body_frixion(Simulation* s, int body, float f) { if (Body* b = body_at(s, body)) b->friction = f; }
float get_body_frixion(Simulation* s, int body) { Body* b = body_at(s, body); return b ? b->friction : 0.0f; }
void set_body_damping(Simulation* s, int body, float linear, float angular) { if (Body* b = body_at(s, body)) { b->linear_damping = linear; b->angular_damping = angular; } }
float get_body_lin_damping(Simulation* s, int body) { Body* b = body_at(s, body); return b ? b->linear_damping : 0.0f; }
void set_body_no_rotation(Simulation*, int) {}  // (the stand-in's boxes never rotate)
void check_collisions(Simulation* s, int body, int* indices, int* size) {
    int n = 0;
    for (int i = 0; i < s->num_bodies; i++) if (i != body && check_collision(s, body, i)) indices[n++] = i;
    *size = n;
}
bool do_collide_except_for(Simulation* s, int body, int exception_id) {
    for (int i = 0; i < s->num_bodies; i++) if (i != body && i != exception_id && check_collision(s, body, i)) return true;
    return false;
}
void hook_set_body_gravity(Simulation* s, int body, glm::vec3 gravity) { if (Body* b = body_at(s, body)) b->gravity = gravity; }
void hook_set_body_no_collision_response(Simulation* s, int body) { if (Body* b = body_at(s, body)) b->collision_response = false; }
int hook_is_grounded(Simulation* s, int body) {
    Body* b = body_at(s, body);
    return b ? (int)(b->position.y - b->half_extents.y <= 0.55f) : 0;
}
void print_resume(const Simulation* s) { std::cout << "host bodies: " << s->num_bodies << std::endl; }

}  // namespace Bullet
}  // namespace Lustrine
