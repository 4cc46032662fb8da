// Lustrine.hpp — the reference's public C++ API (src/Lustrine.hpp:45-144), same names and argument
// meaning; the particle step behind it runs on the GPU through include/lgpu.h.
#pragma once

#include <string>
#include <vector>

#include "Kernels.hpp"
#include "Simulation.hpp"

#ifndef LUSTRINE_EXPORT
#define LUSTRINE_EXPORT __attribute__((visibility("default")))
#endif

namespace Lustrine {

LUSTRINE_EXPORT void init_simulation(const SimulationParameters* parameters, Simulation* simulation,
                                     const std::vector<Grid>& grids_sand_arg, const std::vector<Grid>& grids_solid_arg, int subdivision);
LUSTRINE_EXPORT void init_simulation(const SimulationParameters* parameters, Simulation* simulation,
                                     const std::vector<Grid>& grids_sand_arg, const std::vector<Grid>& grids_solid_arg);
LUSTRINE_EXPORT void init_simulation_extra_parameters(const SimulationParameters* parameters, Simulation* simulation,
                                                      const std::vector<Grid>& grids_sand_arg, const std::vector<Grid>& grids_solid_arg,
                                                      int subdivision, float kernel_radius_scale, bool with_credits);
LUSTRINE_EXPORT void clean_simulation(Simulation* simulation);

LUSTRINE_EXPORT void init_chunk_from_grid(Chunk* chunk, const Grid* grid, MaterialType type, float cell_size, int subdivision, bool mask_out);
LUSTRINE_EXPORT void init_grid_box(const SimulationParameters* parameters, Grid* grid, int X, int Y, int Z, glm::vec3 position, glm::vec4 color, MaterialType type);
LUSTRINE_EXPORT void init_grid_box_random(const SimulationParameters* parameters, Grid* grid, int X, int Y, int Z, glm::vec3 position, glm::vec4 color,
                                          MaterialType type, float probability);
LUSTRINE_EXPORT void init_grid_from_magika_voxel(Grid* grid, const std::string& path, glm::vec3 position, MaterialType type);

LUSTRINE_EXPORT int add_particle_source(Simulation* simulation, const Grid* pattern, glm::vec3 direction, float freq, int capacity);
LUSTRINE_EXPORT int add_particle_sink(Simulation* simulation, glm::vec3 min_pos, glm::vec3 max_pos, float frequency);
LUSTRINE_EXPORT int add_particle_sink(Simulation* simulation, const Grid* pattern, float frequency);  // not implemented in the reference either
LUSTRINE_EXPORT void set_source_state(Simulation* simulation, int index, bool state);
LUSTRINE_EXPORT void set_sink_state(Simulation* simulation, int index, bool state);
LUSTRINE_EXPORT int get_source_spawned(Simulation* simulation, int index);
LUSTRINE_EXPORT int get_sink_despawned(Simulation* simulation, int index);
LUSTRINE_EXPORT int query_cell_num_particles(Simulation* simulation, glm::vec3 min, glm::vec3 max, bool incl_solid);

LUSTRINE_EXPORT void simulate(Simulation* simulation, float dt);

void update_wind_system(WindSystem* wind_system, float dt);

// ---- B200 additions -------------------------------------------------------------------------
namespace B200 {
// How the host arrays of Simulation follow the device state:
//   SYNC_FULL     every simulate_fun call uploads positions/velocities/attracted, steps, downloads them
//                 (exact drop-in: callers may edit the host arrays between calls) — default
//   SYNC_RESIDENT the device state is authoritative; only positions/velocities/attracted are downloaded
//                 after each call (callers read, never write)
//   SYNC_LAZY     nothing is downloaded until sync_to_host() / Wrapper::simulation_bind_positions_copy
enum HostSync { SYNC_FULL = 0, SYNC_RESIDENT = 1, SYNC_LAZY = 2 };
LUSTRINE_EXPORT void set_host_sync(Simulation* simulation, HostSync mode);
LUSTRINE_EXPORT void sync_to_host(Simulation* simulation);
LUSTRINE_EXPORT void copy_positions_to(Simulation* simulation, float* dst);  // 3 * num_sand_particles floats
// Solver options that the reference does not have (defaults reproduce the reference).
LUSTRINE_EXPORT void set_solver_options(Simulation* simulation, int fluid_iterations, bool literal_lambda_index, bool exact_math);
LUSTRINE_EXPORT float last_step_ms(Simulation* simulation);
}  // namespace B200

}  // namespace Lustrine
