// Simulation.hpp — the public state struct of the Lustrine API, restated from the reference's
// src/Simulation.hpp:14-251 (same type names, same field names, same defaults) so that C++ callers
// that read simulation.positions / .colors / .num_sand_particles / .positions_solid directly
// (experiments/fluid/fluid.cpp:85-86,118-120) keep working.  Host arrays keep the reference's
// layout: sand at [0, ptr_sand_end), solids at the tail [total_allocated - num_solid, total_allocated)
// (src/Lustrine.cpp:109-127,232-233).  The particle state itself lives on the GPU (field `gpu`).
#pragma once

#include <cstddef>
#include <utility>
#include <vector>

#ifdef LUSTRINE_B200_BULLET_HEADER
#include LUSTRINE_B200_BULLET_HEADER   // the reference's BulletPhysics.hpp: a real Bullet world on the host (see BulletPhysics.hpp)
#else
#include "BulletPhysics.hpp"
#endif
#include "glm_compat.hpp"

namespace Lustrine {

struct Simulation;
typedef float (*W_fun)(const Simulation*, float);
typedef glm::vec3 (*gradW_fun)(const Simulation*, const glm::vec3&);
typedef void (*Simulate_fun)(Simulation*, float);

enum MaterialType { SOLID = 0, SAND = 1 };

struct Chunk {
    std::vector<glm::vec3> positions;
    std::vector<glm::vec4> colors;
    glm::vec4 color;
    std::vector<int> cells;
    bool has_one_color_per_particles = false;
    int num_particles;
    MaterialType type;
};

struct Grid {
    std::vector<int> cells;
    std::vector<glm::vec4> colors;
    glm::vec4 color;
    bool has_one_color_per_cell;
    int X;
    int Y;
    int Z;
    int num_grid_cells;
    int num_occupied_grid_cells;
    MaterialType type;
    glm::vec3 position;
    bool sparse_solid;
    bool dynamic_solid;
};

struct SimulationParameters {
    int X;
    int Y;
    int Z;
    float particleRadius;
    float particleDiameter;
};

struct ParticleSink {
    std::vector<bool> state;
    std::vector<float> timers;
    std::vector<float> frequencies;
    std::vector<int> despawned;
    std::vector<std::vector<int>> sink_cells;
    int num_sinks = 0;
    std::vector<int> temp_removal;
};

struct ParticleSource {
    std::vector<Chunk> patterns;
    std::vector<glm::vec3> directions;
    std::vector<float> frequencies;
    std::vector<float> timers;
    std::vector<int> capacities;
    std::vector<int> spawned;
    std::vector<bool> source_state;
    int num_sources = 0;
};

struct WindSystem {
    glm::vec3 direction;
    float magnitude;
    float peak;
    float freq_blow;
    float timer_blow;
    float t1;
    float t2;
};

struct CountingSortArrays {
    int* counts = nullptr;
    int* particles_unsorted_indices = nullptr;
    int* particles_sorted_indices = nullptr;
};

struct Simulation {
    Simulate_fun simulate_fun = nullptr;
    Bullet::Simulation bullet_physics_simulation;
    W_fun W = nullptr;
    gradW_fun gradW = nullptr;

    SimulationParameters parameters_copy;
    int subdivision;
    float particleRadius;
    float particleDiameter;
    float kernelRadius;
    float kernelFactor = 0.5f;

    float cubic_kernel_k;
    float cubic_kernel_l;

    float domainX = 30.0f;
    float domainY = 35.0f;
    float domainZ = 30.0f;

    float rest_density = 24.0;
    float mass = 5.0;

    glm::vec3 gravity = glm::vec3(0, -10.0, 0.0);
    float time_step = 0.01f;

    float relaxation_epsilon = 10.0f;

    float s_corr_dq = 0.5f;
    float s_corr_k = 1.0;
    float s_corr_n = 4;

    float c_xsph = 0.1f;
    float epsilon_vorticity = 0.1f;

    glm::vec3* velocities = nullptr;
    std::vector<float> lambdas;
    std::vector<std::vector<int>> neighbors;  // kept for layout; never filled (lists live on the GPU)
    int gridX, gridY, gridZ;
    float cell_size;
    int num_grid_cells;

    std::vector<std::vector<int>> uniform_gird_cells;               // kept for layout; the grid lives on the GPU
    std::vector<std::vector<int>> uniform_grid_cells_static_saved;  // idem
    bool computed_static_particles = false;
    std::vector<std::pair<int, int>> sand_particle_cell_id;
    glm::vec3* velocity_tmp = nullptr;
    glm::vec3* position_star_neighbor_tmp = nullptr;
    glm::vec3* position_neighbor_tmp = nullptr;

    int num_particles = 0;

    size_t total_allocated = 0;
    size_t leftover_allocated = 0;

    glm::vec3* positions = nullptr;
    glm::vec3* positions_star = nullptr;
    glm::vec3* positions_solid = nullptr;

    glm::vec4* colors = nullptr;
    glm::vec4* colors_solid = nullptr;

    int ptr_sand_start = -1;
    int ptr_sand_end = -1;

    int ptr_solid_start = -1;
    int ptr_solid_end = -1;

    int ptr_solid_ordered_start = -1;
    int ptr_solid_ordered_end = -1;

    int num_solid_particles;
    std::vector<Grid> grids_solid;
    std::vector<glm::vec3> grids_initial_positions_solid;
    std::vector<std::pair<int, int>> solid_grid_to_body;
    std::vector<std::pair<int, int>> grids_solid_chunk_ptrs;
    std::vector<Chunk> chunks_solid;

    int num_sand_particles;
    std::vector<Grid> grids_sand;
    std::vector<glm::vec3> grids_initial_positions_sand;
    std::vector<Chunk> chunks_sand;
    int num_remaining_sand_particles;
    glm::vec3* positions_tmp = nullptr;
    bool first_iteration = true;

    bool attract_flag = false;
    bool blow_flag = false;
    int* attracted = nullptr;
    int* attracted_tmp = nullptr;

    float attract_radius = 1.5f;
    float blow_radius = 2.0f;
    float attract_coeff = 1000.0f;
    float blow_coeff = 500.0f;

    ParticleSource* source;
    ParticleSink* sink;
    WindSystem* wind_system;
    CountingSortArrays* counting_sort_arrays;

    float total_time = 0.0f;

    // ---- B200 additions (after every reference field, so the reference prefix keeps its layout) ----
    void* gpu = nullptr;          // B200::DeviceState*
};

}  // namespace Lustrine
