// Lustrine.cpp — lifecycle and per-frame driver of the Lustrine API on top of the GPU step.
//
// Mirrors the behaviour of the reference's src/Lustrine.cpp (citations inline): the host arrays
// keep the reference's layout and contents so existing callers keep working, while the particle
// state that the solver touches lives in an lgpu context (include/lgpu.h).  Sources, sinks and
// query_cell_num_particles — which index the host uniform grid in the reference — go through the
// device grid instead (lgpu_append_sand, lgpu_remove_in_cells, lgpu_cell_count).
#include "lustrine/Lustrine.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <new>

#include "DeviceState.hpp"
#include "lustrine/Profiling.hpp"
#include "lustrine/Simulate.hpp"

namespace Lustrine {

namespace {

constexpr double kPi = 3.14159265358979323846;  // src/Lustrine.cpp:16
const std::align_val_t kAlign{64};              // 64-byte aligned host arrays, src/Lustrine.cpp:129

template <class T> T* aligned_array(size_t n) { return static_cast<T*>(::operator new[](sizeof(T) * (n ? n : 1), kAlign)); }
template <class T> void aligned_free(T* p) { if (p) ::operator delete[](p, kAlign); }

struct CreditsGlyphs;  // the end-credits bitmap font (src/Font.hpp) is game cosmetics and out of scope

void place_particle(Simulation* s, int slot, const Chunk& chunk, int j) {
    s->positions[slot] = chunk.positions[j];
    s->positions_star[slot] = chunk.positions[j];
    s->colors[slot] = chunk.has_one_color_per_particles ? chunk.colors[j] : chunk.color;
}

// The shared body of init_simulation (src/Lustrine.cpp:61-297) and
// init_simulation_extra_parameters (:299-589).
void init_common(const SimulationParameters* parameters, Simulation* s, const std::vector<Grid>& sand_arg,
                 const std::vector<Grid>& solid_arg, int subdivision, float kernel_radius_scale, bool with_credits,
                 bool mask_out_solids) {
    Profiling::init_profiling();
    Bullet::init_bullet(&s->bullet_physics_simulation);
    s->simulate_fun = with_credits ? simulate_sand_credits : simulate_sand;  // :72 / :312 (sand is the default)
    s->subdivision = subdivision;
    s->parameters_copy = *parameters;

    s->source = new ParticleSource();
    s->sink = new ParticleSink();
    s->sink->temp_removal.reserve(100000);
    s->wind_system = new WindSystem();
    s->wind_system->direction = glm::vec3(0, 0, -1);
    s->wind_system->magnitude = 5.0f;
    s->bullet_physics_simulation.particle_radius = parameters->particleRadius;

    s->domainX = parameters->X; s->domainY = parameters->Y; s->domainZ = parameters->Z;  // :105-107

    // capacity, in the reference's int arithmetic (:109-127)
    const int sub3 = subdivision * subdivision * subdivision;
    int first_guess = (int)(s->domainX * s->domainY * s->domainZ * subdivision * subdivision * subdivision);
    int count = 0;
    for (const Grid& g : sand_arg) count += g.num_occupied_grid_cells;
    s->num_sand_particles = count;
    for (const Grid& g : solid_arg) count += g.num_occupied_grid_cells * sub3;
    s->num_solid_particles = count - s->num_sand_particles;
    s->total_allocated = (size_t)std::max(count, first_guess);
    s->leftover_allocated = s->total_allocated - count;

    const size_t total = s->total_allocated;
    s->positions = aligned_array<glm::vec3>(total);
    s->positions_star = aligned_array<glm::vec3>(total);
    s->colors = aligned_array<glm::vec4>(total);
    s->positions_tmp = aligned_array<glm::vec3>(total);
    s->attracted = aligned_array<int>(total);
    s->attracted_tmp = aligned_array<int>(total);
    std::memset(s->attracted, 0, total * sizeof(int));
    std::memset(s->positions, 0, total * sizeof(glm::vec3));
    std::memset(s->positions_star, 0, total * sizeof(glm::vec3));
    std::memset(s->positions_tmp, 0, total * sizeof(glm::vec3));
    std::memset(s->colors, 0, total * sizeof(glm::vec4));

    // sand fills [0, ptr_sand_end) in grid order (:144-172)
    s->ptr_sand_start = 0;
    s->ptr_sand_end = 0;
    s->grids_sand = sand_arg;
    for (const Grid& g : sand_arg) s->grids_initial_positions_sand.push_back(g.position);
    for (size_t i = 0; i < s->grids_sand.size(); i++) {
        Chunk chunk;
        init_chunk_from_grid(&chunk, &s->grids_sand[i], SAND, parameters->particleDiameter, 1, false);
        s->chunks_sand.push_back(chunk);
        for (int j = 0; j < chunk.num_particles; j++) place_particle(s, s->ptr_sand_end++, chunk, j);
    }
    (void)with_credits;  // the credits text particles need the reference's bitmap font (out of scope): no extra particles

    // solids fill the tail downwards, ending at [total - num_solid, total) (:174-235)
    s->ptr_solid_ordered_start = (int)total - 1;
    s->ptr_solid_ordered_end = (int)total - 1;
    s->grids_solid = solid_arg;
    for (const Grid& g : solid_arg) s->grids_initial_positions_solid.push_back(g.position);
    s->solid_grid_to_body.assign(s->grids_solid.size(), std::make_pair(-1, -1));
    s->grids_solid_chunk_ptrs.assign(s->grids_solid.size(), std::make_pair(-1, -1));
    for (size_t i = 0; i < s->grids_solid.size(); i++) {
        Chunk chunk;
        init_chunk_from_grid(&chunk, &s->grids_solid[i], SOLID, parameters->particleDiameter, subdivision, mask_out_solids);
        s->chunks_solid.push_back(chunk);
        s->grids_solid_chunk_ptrs[i].second = s->ptr_solid_ordered_end + 1;
        Chunk boxes;  // one unit rigid box per occupied voxel (:193-198)
        init_chunk_from_grid(&boxes, &s->grids_solid[i], SOLID, 1.0f, 1, false);
        for (int j = 0; j < boxes.num_particles; j++)
            Bullet::add_box(&s->bullet_physics_simulation, boxes.positions[j], s->grids_solid[i].dynamic_solid);
        for (int j = 0; j < chunk.num_particles; j++) place_particle(s, s->ptr_solid_ordered_end--, chunk, j);
        s->grids_solid_chunk_ptrs[i].first = s->ptr_solid_ordered_end;
    }
    s->ptr_solid_start = s->ptr_solid_ordered_end + 1;
    s->ptr_solid_end = s->ptr_solid_ordered_start + 1;
    if (s->ptr_solid_ordered_start == s->ptr_solid_ordered_end) {  // no solids (:219-222)
        s->ptr_solid_start = s->ptr_solid_ordered_start;
        s->ptr_solid_end = s->ptr_solid_ordered_end;
    }
    s->positions_solid = s->positions + total - s->num_solid_particles;
    s->colors_solid = s->colors + total - s->num_solid_particles;
    s->num_remaining_sand_particles = s->ptr_solid_ordered_end - s->ptr_sand_end;  // :237

    const int max_sand = s->ptr_solid_ordered_start;  // :239
    s->velocities = aligned_array<glm::vec3>(max_sand > 0 ? max_sand : 1);
    // The reference clears `count` BYTES here (:242); all velocities start at zero instead.
    std::memset(s->velocities, 0, sizeof(glm::vec3) * (max_sand > 0 ? max_sand : 1));
    s->lambdas.assign(max_sand > 0 ? max_sand : 0, 0.0f);

    s->W = cubic_kernel;  // :247-248
    s->gradW = cubic_kernel_grad;

    s->particleRadius = parameters->particleRadius;
    s->particleDiameter = parameters->particleDiameter;
    s->kernelRadius = kernel_radius_scale * parameters->particleRadius;  // :253 / :545
    s->cell_size = 1.0f * s->kernelRadius;
    float h3 = std::pow(s->kernelRadius, 3);  // promoted to double, rounded once (:257)
    s->cubic_kernel_k = 8.0f / (kPi * h3);
    s->cubic_kernel_l = 48.0f / (kPi * h3);
    s->gridX = (int)(s->domainX / s->cell_size) + 1;
    s->gridY = (int)(s->domainY / s->cell_size) + 1;
    s->gridZ = (int)(s->domainZ / s->cell_size) + 1;
    s->num_grid_cells = s->gridX * s->gridY * s->gridZ;
    // The reference allocates two vector-of-vector grids, neighbour lists and sort scratch on the host
    // here (:267-284); those structures live on the GPU, the host members stay empty.
    s->counting_sort_arrays = new CountingSortArrays();

    Bullet::allocate_particles_colliders(&s->bullet_physics_simulation, s->bullet_physics_simulation.num_particles_allocated, s->particleRadius);
    Bullet::bind_foreign_sand_positions(&s->bullet_physics_simulation, s->positions);

    // device state: constants, solids once, sand
    B200::DeviceState* d = B200::DeviceState::create(s, kernel_radius_scale);
    s->gpu = d;
#ifndef LUSTRINE_B200_BULLET_HEADER
    s->bullet_physics_simulation.gpu = d ? d->ctx : nullptr;
#endif

    std::cout << "registered " << s->num_sand_particles << " sand particles and " << s->num_solid_particles
              << " solid particles (B200 device state: " << (d ? "ok" : "FAILED") << ")\n";
}

}  // namespace

void init_simulation(const SimulationParameters* parameters, Simulation* simulation, const std::vector<Grid>& grids_sand_arg,
                     const std::vector<Grid>& grids_solid_arg) {
    init_simulation(parameters, simulation, grids_sand_arg, grids_solid_arg, 1);
}

void init_simulation(const SimulationParameters* parameters, Simulation* simulation, const std::vector<Grid>& grids_sand_arg,
                     const std::vector<Grid>& grids_solid_arg, int subdivision) {
    // mask_out = true: voxels whose value is exactly 1 get a rigid box but no boundary particle (:188)
    init_common(parameters, simulation, grids_sand_arg, grids_solid_arg, subdivision, 3.1f, false, true);
}

void init_simulation_extra_parameters(const SimulationParameters* parameters, Simulation* simulation, const std::vector<Grid>& grids_sand_arg,
                                      const std::vector<Grid>& grids_solid_arg, int subdivision, float kernel_radius_scale, bool with_credits) {
    init_common(parameters, simulation, grids_sand_arg, grids_solid_arg, subdivision, kernel_radius_scale, with_credits, false);  // :479
}

void clean_simulation(Simulation* s) {  // :591-616
    Bullet::clean_bullet(&s->bullet_physics_simulation);
    B200::DeviceState::destroy(static_cast<B200::DeviceState*>(s->gpu));
    s->gpu = nullptr;
    s->computed_static_particles = false;
    aligned_free(s->positions); aligned_free(s->positions_star); aligned_free(s->colors); aligned_free(s->positions_tmp);
    aligned_free(s->velocities); aligned_free(s->attracted); aligned_free(s->attracted_tmp);
    s->positions = s->positions_star = s->positions_tmp = s->velocities = nullptr;
    s->colors = nullptr; s->attracted = s->attracted_tmp = nullptr;
    delete s->source; delete s->sink; delete s->wind_system; delete s->counting_sort_arrays;
    s->source = nullptr; s->sink = nullptr; s->wind_system = nullptr; s->counting_sort_arrays = nullptr;
}

void init_grid_box(const SimulationParameters*, Grid* grid, int X, int Y, int Z, glm::vec3 position, glm::vec4 color, MaterialType type) {  // :618-639
    grid->type = type;
    grid->num_grid_cells = X * Y * Z;
    grid->num_occupied_grid_cells = grid->num_grid_cells;
    grid->X = X; grid->Y = Y; grid->Z = Z;
    grid->has_one_color_per_cell = false;
    grid->cells.assign(grid->num_grid_cells, 1);
    grid->color = color;
    grid->position = position;
    grid->sparse_solid = true;
    grid->dynamic_solid = false;
}

void init_grid_box_random(const SimulationParameters*, Grid* grid, int X, int Y, int Z, glm::vec3 position, glm::vec4 color, MaterialType type,
                          float probability) {  // :642-674 (every cell is counted as occupied there, too)
    grid->type = type;
    grid->num_grid_cells = X * Y * Z;
    grid->num_occupied_grid_cells = 0;
    grid->X = X; grid->Y = Y; grid->Z = Z;
    grid->has_one_color_per_cell = false;
    grid->cells.assign(grid->num_grid_cells, 1);
    for (int y = 0; y < Y; y++)
        for (int x = 0; x < X; x++)
            for (int z = 0; z < Z; z++) {
                float p = ((float)rand()) / RAND_MAX;
                grid->cells[y * X * Z + x * Z + z] = (p <= probability);
                grid->num_occupied_grid_cells++;
            }
    grid->color = color;
    grid->position = position;
    grid->sparse_solid = true;
    grid->dynamic_solid = false;
}

void init_chunk_from_grid(Chunk* chunk, const Grid* grid, MaterialType type, float cell_size, int subdivision, bool mask_out) {  // :676-722
    chunk->type = type;
    chunk->num_particles = grid->num_occupied_grid_cells * subdivision * subdivision * subdivision;
    chunk->has_one_color_per_particles = grid->has_one_color_per_cell;
    chunk->positions.assign(chunk->num_particles, glm::vec3(0, 0, 0));
    if (chunk->has_one_color_per_particles) chunk->colors.assign(chunk->num_particles, glm::vec4(0, 0, 0, 1));
    else chunk->color = grid->color;
    const glm::vec3 half(cell_size * 0.5f, cell_size * 0.5f, cell_size * 0.5f);
    const int X = grid->X * subdivision, Y = grid->Y * subdivision, Z = grid->Z * subdivision;
    int counter = 0;
    for (int x = 0; x < X; x++)
        for (int y = 0; y < Y; y++)
            for (int z = 0; z < Z; z++) {
                int cell = (x / subdivision) * grid->Y * grid->Z + (y / subdivision) * grid->Z + (z / subdivision);
                int value = grid->cells[cell];
                if (!value || (mask_out && value == 1)) continue;
                if (counter >= chunk->num_particles) continue;  // sparse grids whose count was under-reported
                glm::vec3 p(x * cell_size, y * cell_size, z * cell_size);
                p += grid->position;
                p += half;
                chunk->positions[counter] = p;
                if (chunk->has_one_color_per_particles) chunk->colors[counter] = grid->colors[cell];
                counter++;
            }
}

// ---------------------------------------------------------------------------------------------
// per-frame driver: sources -> simulate_fun -> sinks (src/Lustrine.cpp:730-861)
// ---------------------------------------------------------------------------------------------
void simulate(Simulation* s, float dt) {
    B200::DeviceState* d = static_cast<B200::DeviceState*>(s->gpu);
    const float dt_clamped = 0.016f;  // :744
    for (int i = 0; i < s->source->num_sources; i++) s->source->timers[i] += dt_clamped;

    for (int i = 0; i < s->source->num_sources; i++) {  // :749-800
        ParticleSource& src = *s->source;
        if (!(src.source_state[i] && src.spawned[i] < src.capacities[i] && src.timers[i] >= src.frequencies[i])) continue;
        Chunk& pattern = src.patterns[i];
        if (s->num_remaining_sand_particles < pattern.num_particles) continue;
        const float freq = src.frequencies[i];
        const float t_last = s->total_time;  // the reference's `total_time; - timers[i];` is two statements (:757)
        const float offset = freq - fmodf(t_last, freq);
        const float diameter = 2.0f * s->particleRadius;
        const float speed = freq > 0.0f ? 1.0f * diameter / freq : 1.0f;
        const glm::vec3& direction = src.directions[i];
        const glm::vec3 shift = direction * offset;
        const int first = s->ptr_sand_end;
        for (int k = 0; k < pattern.num_particles; k++) {
            s->positions[first + k] = pattern.positions[k] + shift;
            s->velocities[first + k] = direction * speed + s->gravity * 0.0f;
            s->colors[first + k] = pattern.color;
        }
        s->ptr_sand_end += pattern.num_particles;
        s->num_remaining_sand_particles -= pattern.num_particles;
        s->num_sand_particles += pattern.num_particles;
        s->velocities[first + pattern.num_particles / 2] += 0.001f * glm::vec3(1.0f);  // :796
        src.timers[i] = 0.0f;
        src.spawned[i] += pattern.num_particles;
        if (d) d->append_from_host(s, first, pattern.num_particles);
    }

    auto t0 = std::chrono::steady_clock::now();
    s->simulate_fun(s, dt);  // :802-804
    Profiling::record(0, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());

    // sinks: every particle in a sink cell is evicted with the reference's swap-with-last order (:806-836)
    std::vector<int> cells;
    for (int i = 0; i < s->sink->num_sinks; i++) {
        if (s->sink->state[i] && s->sink->timers[i] >= s->sink->frequencies[i])
            cells.insert(cells.end(), s->sink->sink_cells[i].begin(), s->sink->sink_cells[i].end());
    }
    if (!cells.empty() && d) {
        int removed = d->remove_in_cells(s, cells);
        s->ptr_sand_end -= removed;
        s->num_sand_particles -= removed;
        s->num_remaining_sand_particles += removed;
    }
    for (int i = 0; i < s->sink->num_sinks; i++) s->sink->timers[i] += dt;  // :856-858
    s->total_time += dt_clamped;
}

namespace {
struct CellBox { int lo[3], hi[3]; };
// get_cell_indices (src/neighbors/Utils.hpp:13-21): clamp to [cs/2, D - cs/2], divide, truncate
CellBox cell_box(const Simulation* s, glm::vec3 a, glm::vec3 b) {
    auto idx = [&](glm::vec3 p, int* out) {
        const float half = s->cell_size * 0.5f;
        const float D[3] = {s->domainX, s->domainY, s->domainZ};
        for (int k = 0; k < 3; k++) {
            float v = std::min(std::max(p[k], half), D[k] - half);
            out[k] = (int)(v / s->cell_size);
        }
    };
    CellBox r;
    idx(a, r.lo);
    idx(b, r.hi);
    return r;
}
}  // namespace

int query_cell_num_particles(Simulation* s, glm::vec3 min_pos, glm::vec3 max_pos, bool include_solid) {  // :864-914
    B200::DeviceState* d = static_cast<B200::DeviceState*>(s->gpu);
    if (!d) return 0;
    CellBox b = cell_box(s, min_pos, max_pos);
    return d->cell_count(b.lo, b.hi, include_solid);
}

int add_particle_source(Simulation* s, const Grid* pattern, glm::vec3 direction, float freq, int capacity) {  // :917-935
    Chunk chunk;
    init_chunk_from_grid(&chunk, pattern, SAND, s->particleDiameter, 1, false);
    int index = s->source->num_sources++;
    s->source->patterns.push_back(chunk);
    direction /= glm::length(direction);
    s->source->directions.push_back(direction);
    s->source->timers.push_back(0.0f);
    s->source->frequencies.push_back(freq);
    s->source->source_state.push_back(true);
    s->source->capacities.push_back(capacity < 0 ? std::numeric_limits<int>::max() : capacity);
    s->source->spawned.push_back(0);
    return index;
}

int add_particle_sink(Simulation* s, glm::vec3 min_pos, glm::vec3 max_pos, float frequency) {  // :937-974
    CellBox b = cell_box(s, min_pos, max_pos);
    int index = s->sink->num_sinks++;
    s->sink->sink_cells.emplace_back();
    s->sink->despawned.push_back(0);
    s->sink->frequencies.push_back(frequency);
    s->sink->state.push_back(true);
    s->sink->timers.push_back(0.0f);
    std::vector<int>& cells = s->sink->sink_cells[index];
    for (int y = b.lo[1]; y <= b.hi[1]; y++)
        for (int x = b.lo[0]; x <= b.hi[0]; x++)
            for (int z = b.lo[2]; z <= b.hi[2]; z++) cells.push_back(y * s->gridX * s->gridZ + x * s->gridZ + z);
    std::sort(cells.begin(), cells.end());
    return index;
}

int add_particle_sink(Simulation*, const Grid*, float) {  // :978-1001: announced as not implemented in the reference
    std::cout << "sink add with grid not implemented yet!" << std::endl;
    return 0;
}

void set_source_state(Simulation* s, int index, bool state) { s->source->source_state[index] = state; }
void set_sink_state(Simulation* s, int index, bool state) { s->sink->state[index] = state; }
int get_source_spawned(Simulation* s, int index) { return s->source->spawned[index]; }
int get_sink_despawned(Simulation* s, int index) { return s->sink->despawned[index]; }
void update_wind_system(WindSystem*, float) {}

namespace B200 {
void set_host_sync(Simulation* s, HostSync mode) { if (s->gpu) static_cast<DeviceState*>(s->gpu)->sync_mode = mode; }
void sync_to_host(Simulation* s) { if (s->gpu) static_cast<DeviceState*>(s->gpu)->download(s); }
void copy_positions_to(Simulation* s, float* dst) {
    DeviceState* d = static_cast<DeviceState*>(s->gpu);
    // (the device holds the same positions as the host arrays right after a step: a page-locked destination gets them
    // from the copy engine — faster than a host memcpy of 12 B per particle)
    if (d && (d->sync_mode == SYNC_LAZY || (d->device_matches_host && lgpu_host_is_pinned(dst)))) d->download_positions_into(dst);
    else std::memcpy(dst, s->positions, sizeof(float) * 3 * s->num_sand_particles);
}
void set_solver_options(Simulation* s, int fluid_iterations, bool literal_lambda_index, bool exact_math) {
    DeviceState* d = static_cast<DeviceState*>(s->gpu);
    if (!d) return;
    d->fluid_iterations = fluid_iterations < 1 ? 1 : fluid_iterations;
    d->literal_lambda_index = literal_lambda_index;
    d->exact_math = exact_math;
}
float last_step_ms(Simulation* s) { return s->gpu ? static_cast<DeviceState*>(s->gpu)->last_ms : 0.0f; }
}  // namespace B200

}  // namespace Lustrine
