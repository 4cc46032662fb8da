// plugin_hook.cpp — TEST INFRASTRUCTURE: the reference-side binding of INTEGRATION.md §A, compiled INSIDE the reference's
// build (its headers, its objects) and linked with liblgpu.so: two `Simulate_fun`s (src/Simulation.hpp:22) that hand the
// particle step to the device boundary of include/lgpu.h and leave everything else of the reference as it is —
// Lustrine::simulate, the host arrays, Bullet.  tests/test_plugin_hook.py runs the reference with its stock
// simulate_sand / fluid step and with these, frame by frame.
//
// The reference's Simulation struct has no member for a device context (INTEGRATION.md suggests adding `void* gpu`);
// the test keeps the context of the one simulation it runs in a static instead, so that the reference sources stay
// unmodified.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "Lustrine.hpp"
#include "Simulate.hpp"
#include "BulletPhysics.hpp"
#include "lgpu.h"

// oracle/ref_harness.cpp: the reference's fluid loop with the delta-p output double-buffered
extern "C" void refh_use_jacobi(void* simulation, int iterations, int literal_lambda_index);

namespace Lustrine {

static lgpu_ctx* g_ctx = nullptr;

static void must(int status, const char* what) {
    if (status != LGPU_OK) { std::fprintf(stderr, "%s: %s\n", what, lgpu_last_error()); std::abort(); }
}

static lgpu_ctx* ctx_of(Simulation* s) {
    if (g_ctx) return g_ctx;
    lgpu_config cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.domain[0] = s->parameters_copy.X; cfg.domain[1] = s->parameters_copy.Y; cfg.domain[2] = s->parameters_copy.Z;
    cfg.particle_radius = s->particleRadius;
    cfg.particle_diameter = s->particleDiameter;
    cfg.kernel_radius_scale = s->kernelRadius / s->particleRadius;  // 3.1f unless init_simulation_extra_parameters
    cfg.capacity_sand = s->ptr_sand_end + 1024;
    cfg.capacity_solid = s->num_solid_particles;
    cfg.device = -1;
    must(lgpu_create(&cfg, &g_ctx), "lgpu_create");
    if (s->num_solid_particles) must(lgpu_upload_solids(g_ctx, s->num_solid_particles, &s->positions[s->ptr_solid_start].x), "lgpu_upload_solids");
    return g_ctx;
}

static void common_params(Simulation* s, float dt, lgpu_step_params* p) {
    lgpu_default_step_params(p);
    p->dt = dt; p->exact_math = 1;
    p->gravity[0] = s->gravity.x; p->gravity[1] = s->gravity.y; p->gravity[2] = s->gravity.z;
    p->mass = s->mass;
}

void simulate_sand_b200(Simulation* s, float dt) {
    Bullet::simulate_bullet(&s->bullet_physics_simulation, dt, s->ptr_sand_start, s->ptr_sand_end);  // host, as src/Simulate.cpp:157
    lgpu_ctx* c = ctx_of(s);
    const int n = s->ptr_sand_end - s->ptr_sand_start;
    must(lgpu_upload_sand(c, n, &s->positions[0].x, &s->velocities[0].x, s->attracted), "lgpu_upload_sand");
    lgpu_step_params p;
    common_params(s, dt, &p);
    p.iterations = 4;
    const glm::vec3& pp = s->bullet_physics_simulation.player_position;
    p.player_position[0] = pp.x; p.player_position[1] = pp.y; p.player_position[2] = pp.z;
    static bool prev_attract = false;  // as src/Simulate.cpp:185
    p.attract_flag = s->attract_flag; p.blow_flag = s->blow_flag; p.prev_attract_flag = prev_attract;
    p.attract_radius = s->attract_radius; p.blow_radius = s->blow_radius;
    p.attract_coeff = s->attract_coeff; p.blow_coeff = s->blow_coeff;
    must(lgpu_step_sand(c, &p), "lgpu_step_sand");
    must(lgpu_download_sand(c, &s->positions[0].x, &s->velocities[0].x, s->attracted), "lgpu_download_sand");
    std::memcpy(s->positions_star, s->positions, sizeof(glm::vec3) * n);  // src/Simulate.cpp:318
    prev_attract = s->attract_flag; s->time_step = dt; s->first_iteration = false;
}

void simulate_fluid_b200(Simulation* s, float dt) {
    Bullet::simulate_bullet(&s->bullet_physics_simulation, dt, s->ptr_sand_start, s->ptr_sand_end);  // src/Simulate.cpp:29
    lgpu_ctx* c = ctx_of(s);
    const int n = s->ptr_sand_end - s->ptr_sand_start;
    must(lgpu_upload_sand(c, n, &s->positions[0].x, &s->velocities[0].x, nullptr), "lgpu_upload_sand");
    lgpu_step_params p;
    common_params(s, dt, &p);
    p.iterations = 1; p.literal_lambda_index = 1;  // the loop as shipped, src/Simulate.cpp:27-115
    p.rest_density = s->rest_density; p.relaxation_epsilon = s->relaxation_epsilon;
    p.s_corr_dq = s->s_corr_dq; p.s_corr_k = s->s_corr_k; p.s_corr_n = s->s_corr_n;
    must(lgpu_step_fluid(c, &p), "lgpu_step_fluid");
    must(lgpu_download_sand(c, &s->positions[0].x, &s->velocities[0].x, nullptr), "lgpu_download_sand");
    std::memcpy(s->positions_star, s->positions, sizeof(glm::vec3) * n);
    s->time_step = dt;
}

}  // namespace Lustrine

// which: 0 stock simulate_sand, 1 simulate_sand_b200, 2 the reference's fluid loop (double-buffered, K = 1, literal index),
// 3 simulate_fluid_b200.  A sand block of side^3 over a one-cell floor slab; out = 3 * side^3 floats per recorded frame.
extern "C" __attribute__((visibility("default"))) int plugin_run(int which, int side, int steps, float* out, int* n_out) {
    using namespace Lustrine;
    Simulation simulation;
    SimulationParameters parameters;
    parameters.X = 3 * side + 6; parameters.Y = 2 * side + 8; parameters.Z = 3 * side + 6;
    parameters.particleRadius = 0.5f; parameters.particleDiameter = 1.0f;
    std::vector<Grid> sand(1), solids(1);
    init_grid_box(&parameters, &sand[0], side, side, side, glm::vec3(side + 2.0f, 3.0f, side + 2.0f), glm::vec4(1.0f), MaterialType::SAND);
    init_grid_box(&parameters, &solids[0], 3 * side + 4, 1, 3 * side + 4, glm::vec3(0.0f), glm::vec4(0.5f), MaterialType::SOLID);
    init_simulation(&parameters, &simulation, sand, solids);
    for (int i = 0; i < simulation.ptr_sand_end; i++) simulation.velocities[i] = glm::vec3(0.3f * ((i % 7) - 3), 0.0f, 0.2f * ((i % 5) - 2));  // 
    switch (which) {
        case 0: simulation.simulate_fun = simulate_sand; break;
        case 1: simulation.simulate_fun = simulate_sand_b200; break;
        case 2: refh_use_jacobi(&simulation, 1, 1); break;
        default: simulation.simulate_fun = simulate_fluid_b200; break;
    }
    const int n = simulation.ptr_sand_end - simulation.ptr_sand_start;
    *n_out = n;
    for (int k = 0; k < steps; k++) {
        simulate(&simulation, which < 2 ? 0.016f : 0.01f);  // Lustrine::simulate: the reference's frame function calls the hook
        std::memcpy(out + (size_t)k * 3 * n, simulation.positions, sizeof(float) * 3 * n);
    }
    clean_simulation(&simulation);
    if (Lustrine::g_ctx) { lgpu_destroy(Lustrine::g_ctx); Lustrine::g_ctx = nullptr; }
    return 0;
}
