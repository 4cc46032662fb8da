// BulletPhysics.hpp — host-side rigid-body hook of the particle step.
//
// In the reference, Bullet3 rigid bodies live entirely on the host (src/BulletPhysics.cpp) and
// touch the particle step at exactly two points:
//   * simulate_bullet() runs first in every simulate_* (src/Simulate.cpp:29,157) and publishes
//     Bullet::Simulation::player_position, which the sand predict reads for attract / blow;
//   * set_particles_box_colliders_positions() scans ALL sand positions for the first <= 100
//     particles inside the player's AABB and parks kinematic proxy boxes on them
//     (src/BulletPhysics.cpp:602-652).
// BASELINE.json keeps Bullet on the host, and the Bullet glue itself is outside the hot-path
// scope.  This header keeps the reference's Bullet::Simulation fields
// that the particle step and the C wrapper touch and the same function names; the bodies are
// stored in a small host-side table (kinematic: gravity + explicit Euler for dynamic bodies,
// no contact solver).  The O(N) host scan is replaced by the device compaction
// lgpu_aabb_first_k.  To couple a real Bullet world, build the host library with
// -DLUSTRINE_B200_BULLET_HEADER='"<reference>/src/BulletPhysics.hpp"': Simulation.hpp then embeds the reference's
// Bullet::Simulation, HostBodies.cpp is left out and the reference's own BulletPhysics.cpp + Bullet are linked instead
// (tests/cpp/Makefile builds exactly that for the parity test of the mixed scene).
#pragma once

#include <vector>

#include "glm_compat.hpp"

namespace Lustrine {
namespace Bullet {

struct Body {
    glm::vec3 position = glm::vec3(0.0f);
    glm::vec3 velocity = glm::vec3(0.0f);
    glm::vec3 half_extents = glm::vec3(0.5f);
    glm::vec3 gravity = glm::vec3(0.0f, -25.0f, 0.0f);
    bool dynamic = false;
    bool detector = false;
    bool collision_response = true;
    float friction = 0.0f, linear_damping = 0.0f, angular_damping = 0.0f;
};

struct Simulation {
    glm::vec3 gravity = glm::vec3(0.0f, -25.0f, 0.0f);  // src/BulletPhysicsSimulation.hpp:23
    int num_bodies = 0;

    int num_particles_allocated = 100;  // kinematic proxy boxes around the player (:29)
    int player_id = 0;
    glm::vec3 player_position = glm::vec3(0.0f);
    float player_box_radius = 4.0f;
    glm::vec3 player_box_scale = glm::vec3(0.0f);
    float particle_radius = 0.1f;

    bool allocated_particles_bounding_boxes = false;
    bool particles_bounding_box_requested_state = true;
    bool particles_bounding_box_current_state = true;

    glm::vec3* foreign_sand_positions = nullptr;  // not owned
    size_t ptr_bounding_box_start = 0;
    size_t ptr_bounding_box_end = 0;
    bool bounding_box_activated = true;

    std::vector<Body> bodies;
    std::vector<std::vector<int>> bodies_collisions;
    void* gpu = nullptr;  // lgpu_ctx* used for the player-AABB compaction (not owned)
};

void init_bullet(Simulation* simulation);
void clean_bullet(Simulation* simulation);
void simulate_bullet(Simulation* simulation, float dt, int sand_start, int sand_end);
void set_gravity(Simulation* simulation, glm::vec3 new_gravity);
glm::vec3 get_gravity(Simulation* simulation);
int add_box(Simulation* simulation, glm::vec3 position, bool is_dynamic);
int add_box(Simulation* simulation, glm::vec3 position, bool is_dynamic, glm::vec3 half_dims);
int add_capsule(Simulation* simulation, glm::vec3 position, float radius, float height);
int add_detector_block(Simulation* simulation, glm::vec3 position, glm::vec3 half_dims);
void allocate_particles_colliders(Simulation* simulation, int num_particles, float radius);
void bind_foreign_sand_positions(Simulation* simulation, glm::vec3* positions);
void set_particles_box_colliders_positions(Simulation* simulation, glm::vec3* particles, int start_ptr, int end_ptr);
void enable_particles_bounding_boxes(Simulation* simulation);
void disable_particles_bounding_boxes(Simulation* simulation);
bool check_collision(Simulation* simulation, int body1, int body2);
bool do_collide(Simulation* simulation, int body);
int get_num_bodies(Simulation* simulation);
void apply_impulse(Simulation* simulation, int body, glm::vec3 impulse, glm::vec3 relative_pos);
glm::vec3 get_body_position(Simulation* simulation, int body);
glm::vec3 get_body_velocity(Simulation* simulation, int body);
void set_body_position(Simulation* simulation, int body, glm::vec3 position);
void set_body_velocity(Simulation* simulation, int body, glm::vec3 velocity);
void add_body_velocity(Simulation* simulation, int body, glm::vec3 velocity);
void set_body_frixion(Simulation* simulation, int body, float frixion);           // (sic) src/BulletPhysics.hpp:59-63
float get_body_frixion(Simulation* simulation, int body);
void set_body_damping(Simulation* simulation, int body, float linear, float angular);
float get_body_lin_damping(Simulation* simulation, int body);
void set_body_no_rotation(Simulation* simulation, int body_index);
void check_collisions(Simulation* simulation, int body, int* indices, int* size);  // :72-74
bool do_collide_except_for(Simulation* simulation, int body, int exception_id);
void print_resume(const Simulation* simulation);

}  // namespace Bullet
}  // namespace Lustrine
