// DeviceState.cpp — glue between the Lustrine::Simulation host struct and the lgpu C ABI.
#include "DeviceState.hpp"

#include <cstdlib>
#include <cstring>
#include <iostream>

#include "lustrine/Profiling.hpp"

namespace Lustrine {
namespace B200 {

namespace {
void die(const char* what, int status) {
    // The particle step has no CPU fallback: a failing device call is fatal and loud.
    std::cerr << "lustrine_b200: " << what << " failed (status " << status << "): " << lgpu_last_error() << std::endl;
    std::abort();
}
#define LGPU_MUST(call) do { int _s = (call); if (_s != LGPU_OK) die(#call, _s); } while (0)
}  // namespace

// The device context is sized for the particles that exist, with head room, and GROWS when the particle sources outgrow
// it — not for every slot of the reference's host arrays (total_allocated = X*Y*Z slots, src/Lustrine.cpp:237-239: 12 M
// slots for the 1 M dam break, 192 M for the 16 M domain).  LUSTRINE_B200_MAX_SAND pins the capacity instead.
void DeviceState::make_context(Simulation* s, int capacity_sand) {
    lgpu_config cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.domain[0] = s->parameters_copy.X; cfg.domain[1] = s->parameters_copy.Y; cfg.domain[2] = s->parameters_copy.Z;
    cfg.particle_radius = s->parameters_copy.particleRadius;
    cfg.particle_diameter = s->parameters_copy.particleDiameter;
    cfg.kernel_radius_scale = kernel_radius_scale;
    cfg.capacity_sand = capacity_sand;
    cfg.capacity_solid = s->num_solid_particles;
    cfg.device = -1;
    LGPU_MUST(lgpu_create(&cfg, &ctx));
    capacity = capacity_sand;
#ifndef LUSTRINE_B200_BULLET_HEADER
    s->bullet_physics_simulation.gpu = ctx;  // (the rigid-body side borrows the context for the player-AABB compaction)
#endif
    lgpu_grid_info gi;
    LGPU_MUST(lgpu_get_grid(ctx, &gi));
    if (gi.grid[0] != s->gridX || gi.grid[1] != s->gridY || gi.grid[2] != s->gridZ || gi.kernel_radius != s->kernelRadius ||
        gi.cubic_k != s->cubic_kernel_k || gi.cubic_l != s->cubic_kernel_l) {
        std::cerr << "lustrine_b200: device grid constants differ from the host's" << std::endl;
        std::abort();
    }
    if (s->num_solid_particles > 0)
        LGPU_MUST(lgpu_upload_solids(ctx, s->num_solid_particles, reinterpret_cast<const float*>(s->positions + s->ptr_solid_start)));
}

// Page-locks the live part of the host arrays a SYNC_FULL / SYNC_RESIDENT step transfers (positions, positions_star,
// velocities, attracted: the first `capacity` sand slots), so that the copies run at the PCIe rate; the arrays stay the
// caller's (src/Lustrine.cpp:129-141 allocates them, clean_simulation frees them).  LUSTRINE_B200_PIN_HOST=0 keeps them
// pageable; a failing registration is not an error (the transfers work either way).
void DeviceState::pin_host(Simulation* s) {
    unpin_host();
    const char* env = std::getenv("LUSTRINE_B200_PIN_HOST");
    if (env && std::atoi(env) == 0) return;
    const size_t n = (size_t)capacity;
    void* ptrs[4] = {s->positions + s->ptr_sand_start, s->positions_star + s->ptr_sand_start, s->velocities + s->ptr_sand_start,
                     s->attracted + s->ptr_sand_start};
    const size_t bytes[4] = {n * sizeof(glm::vec3), n * sizeof(glm::vec3), n * sizeof(glm::vec3), n * sizeof(int)};
    for (int k = 0; k < 4; k++) {
        if (!ptrs[k] || lgpu_host_register(ptrs[k], bytes[k]) != LGPU_OK) { unpin_host(); return; }
        pinned[k] = ptrs[k];
    }
    host_pinned = true;
}

void DeviceState::unpin_host() {
    for (int k = 0; k < 4; k++) { if (pinned[k]) lgpu_host_unregister(pinned[k]); pinned[k] = nullptr; }
    host_pinned = false;
}

DeviceState* DeviceState::create(Simulation* s, float kernel_radius_scale) {
    DeviceState* d = new DeviceState();
    d->kernel_radius_scale = kernel_radius_scale;
    d->capacity_limit = s->ptr_solid_ordered_end + 1 > 0 ? s->ptr_solid_ordered_end + 1 : 1;
    const int live = s->ptr_sand_end - s->ptr_sand_start;
    int cap = live + live / 2 + 65536;  // head room for the sources; grows on demand
    const char* cap_env = std::getenv("LUSTRINE_B200_MAX_SAND");
    if (cap_env && std::atoi(cap_env) >= live) cap = std::atoi(cap_env);
    if (cap > d->capacity_limit) cap = d->capacity_limit;
    if (cap < live) cap = live;
    d->make_context(s, cap > 0 ? cap : 1);
    d->pin_host(s);
    d->upload(s);
    return d;
}

void DeviceState::ensure_capacity(Simulation* s, int needed) {
    if (needed <= capacity) return;
    // the device state is authoritative unless every step uploads it anyway: bring it home before the context goes
    if (sync_mode != SYNC_FULL) download(s);
    long grown = (long)needed + needed / 2 + 65536;
    if (grown > capacity_limit) grown = capacity_limit;
    if (grown < needed) grown = needed;
    lgpu_destroy(ctx);
    ctx = nullptr;
    make_context(s, (int)grown);
    pin_host(s);
    device_sand = 0;
}

void DeviceState::destroy(DeviceState* d) {
    if (!d) return;
    d->unpin_host();
    lgpu_destroy(d->ctx);
    delete d;
}

// with_flags = false: the step about to run is simulate_fluid, which neither reads nor writes Simulation::attracted
// (src/Simulate.cpp:27-115): 4 bytes per particle less in either direction
void DeviceState::upload(Simulation* s, bool with_flags) {
    const int n = s->ptr_sand_end - s->ptr_sand_start;
    ensure_capacity(s, n);
    LGPU_MUST(lgpu_upload_sand(ctx, n, reinterpret_cast<const float*>(s->positions + s->ptr_sand_start),
                               reinterpret_cast<const float*>(s->velocities + s->ptr_sand_start),
                               with_flags ? s->attracted + s->ptr_sand_start : nullptr));
    device_sand = n;
    device_flags_valid = with_flags;
}

void DeviceState::download(Simulation* s, bool with_flags) {
    // the reference leaves positions_star == positions after a step (src/Simulate.cpp:111,318): with page-locked host
    // arrays the copy engine delivers the positions to both (12 B per particle over PCIe instead of a host memcpy)
    float* star = reinterpret_cast<float*>(s->positions_star + s->ptr_sand_start);
    LGPU_MUST(lgpu_download_sand2(ctx, reinterpret_cast<float*>(s->positions + s->ptr_sand_start), host_pinned ? star : nullptr,
                                  reinterpret_cast<float*>(s->velocities + s->ptr_sand_start),
                                  with_flags && device_flags_valid ? s->attracted + s->ptr_sand_start : nullptr));
    if (!host_pinned) std::memcpy(star, s->positions + s->ptr_sand_start, sizeof(glm::vec3) * (size_t)lgpu_num_sand(ctx));
    device_matches_host = true;
}

void DeviceState::download_positions_into(float* dst) { LGPU_MUST(lgpu_download_sand(ctx, dst, nullptr, nullptr)); }

void DeviceState::append_from_host(Simulation* s, int first, int count) {
    if (sync_mode == SYNC_FULL) { device_matches_host = false; return; }  // the next step uploads everything anyway
    if (device_sand + count > capacity) {
        // the context is re-created larger; `first .. first + count` are already in the host arrays, so one upload brings
        // everything (the old particles just downloaded, and the new ones)
        ensure_capacity(s, device_sand + count);
        upload(s);
        return;
    }
    LGPU_MUST(lgpu_append_sand(ctx, count, reinterpret_cast<const float*>(s->positions + first),
                               reinterpret_cast<const float*>(s->velocities + first), s->attracted + first));
    device_sand += count;
}

int DeviceState::remove_in_cells(Simulation* s, const std::vector<int>& cells) {
    int removed = 0;
    LGPU_MUST(lgpu_remove_in_cells(ctx, cells.data(), (int)cells.size(), &removed));
    device_sand -= removed;
    if (removed > 0 && sync_mode != SYNC_LAZY) download(s);
    return removed;
}

int DeviceState::cell_count(const int lo[3], const int hi[3], bool include_solid) {
    int count = 0;
    LGPU_MUST(lgpu_cell_count(ctx, lo, hi, include_solid ? 1 : 0, &count));
    return count;
}

void DeviceState::step(Simulation* s, float dt, int mode) {
    // (a fluid step leaves `attracted` alone: under SYNC_FULL it is not even uploaded; the device copy of the other
    // modes stays what the last sand step or upload made it)
    // (... unless a particle sink is configured: an eviction moves the last particle's `attracted` along with it,
    // src/Lustrine.cpp:829, and the device does that move)
    const bool with_flags = mode != 1 || (s->sink && s->sink->num_sinks > 0);
    if (sync_mode == SYNC_FULL) upload(s, with_flags);
    else if (with_flags && !device_flags_valid) upload(s, true);  // (a sand step after fluid steps that never brought the flags)
    lgpu_step_params p;
    lgpu_default_step_params(&p);
    p.dt = dt;
    p.gravity[0] = s->gravity.x; p.gravity[1] = s->gravity.y; p.gravity[2] = s->gravity.z;
    p.rest_density = s->rest_density; p.mass = s->mass; p.relaxation_epsilon = s->relaxation_epsilon;
    p.s_corr_dq = s->s_corr_dq; p.s_corr_k = s->s_corr_k; p.s_corr_n = s->s_corr_n;
    p.exact_math = exact_math ? 1 : 0;
    p.sph_kernel = (s->W == static_cast<W_fun>(poly6_kernel)) ? 1 : 0;  // which kernel the caller wired
    if (mode == 1) {
        p.iterations = fluid_iterations;
        p.literal_lambda_index = literal_lambda_index ? 1 : 0;
        LGPU_MUST(lgpu_step_fluid(ctx, &p));
        s->time_step = dt < 0.001f ? 0.001f : (dt > 0.01f ? 0.01f : dt);  // src/Simulate.cpp:31-32
    } else {
        const glm::vec3& pp = s->bullet_physics_simulation.player_position;
        p.iterations = 4;  // src/Simulate.cpp:226
        p.player_position[0] = pp.x; p.player_position[1] = pp.y; p.player_position[2] = pp.z;
        p.attract_flag = s->attract_flag; p.blow_flag = s->blow_flag; p.prev_attract_flag = prev_attract_flag;
        p.attract_radius = s->attract_radius; p.blow_radius = s->blow_radius;
        p.attract_coeff = s->attract_coeff; p.blow_coeff = s->blow_coeff;
        p.credits = mode == 3;
        if (mode == 3) { p.mu_s = 0.8f; p.mu_k = 0.7f; }  // src/Simulate.cpp:333-334
        LGPU_MUST(lgpu_step_sand(ctx, &p));
        s->time_step = dt;  // :168
        prev_attract_flag = s->attract_flag;  // :323
        s->first_iteration = false;
    }
    if (sync_mode != SYNC_LAZY) download(s, with_flags);
    else { device_matches_host = false; LGPU_MUST(lgpu_sync(ctx)); }
    LGPU_MUST(lgpu_last_step_ms(ctx, 0, &last_ms));
    Profiling::record(2, last_ms * 1e-3);
}

}  // namespace B200
}  // namespace Lustrine
