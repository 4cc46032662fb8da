// LustrineWrapper.hpp — the C entry points of the reference's DLL (src/LustrineWrapper.hpp:30-188):
// same unmangled names, same POD structs, one process-global simulation.  A game engine binding
// written against the reference's LustrineWrapper keeps working against liblustrine_b200.so.
#pragma once

#include <cstddef>
#include <cstdint>

#include "Lustrine.hpp"

#ifndef LUSTRINE_WRAPPER_EXPORT
#define LUSTRINE_WRAPPER_EXPORT __attribute__((visibility("default")))
#endif

namespace Lustrine {
namespace Wrapper {

struct SimulationData {
    int num_sand_particles;
    int num_solid_particles;
    int start_sand_index;
    int end_sand_index;
    int start_solid_index;
    int end_solid_index;
};

struct Color { float r, g, b, a; };
struct Vec3 { float x, y, z; };

struct GridWrapper {
    int* cells;
    Color* colors;
    Color color;
    Vec3 position;
    bool has_one_color_per_cell;
    int X;
    int Y;
    int Z;
    int num_grid_cells;
    int num_occupied_grid_cells;
    int type;
};

struct BindingString {
    int64_t length;
    char* data;
};

extern "C" {
LUSTRINE_WRAPPER_EXPORT void init_simulation(const SimulationParameters* parameters, SimulationData* data, GridWrapper* sand_grids,
                                             int num_sand_grids, GridWrapper* solid_grids, int num_solid_grids, int subdivision);
LUSTRINE_WRAPPER_EXPORT void init_simulation_extra_parameters(const SimulationParameters* parameters, SimulationData* data, GridWrapper* sand_grids,
                                                              int num_sand_grids, GridWrapper* solid_grids, int num_solid_grids, int subdivision,
                                                              float kernel_radius_scale, int with_credits);
LUSTRINE_WRAPPER_EXPORT void simulate(float dt, bool attract_flag, bool blow_flag);
LUSTRINE_WRAPPER_EXPORT void simulate_no_flags(float dt);
LUSTRINE_WRAPPER_EXPORT void simulation_bind_positions_copy(float* position_ptr);
LUSTRINE_WRAPPER_EXPORT void cleanup_simulation();
LUSTRINE_WRAPPER_EXPORT void init_grid_box(const SimulationParameters* parameters, GridWrapper* grid, int X, int Y, int Z, Vec3 position, Color color, int type);
LUSTRINE_WRAPPER_EXPORT void read_vox_scene(BindingString* data, const uint8_t* buffer, int64_t size);
LUSTRINE_WRAPPER_EXPORT void free_string(BindingString* data);
LUSTRINE_WRAPPER_EXPORT void create_grid(GridWrapper* grid, const wchar_t* path, int type, int pathlen);
LUSTRINE_WRAPPER_EXPORT void init_grid_magikavoxel(GridWrapper* grid, const char* path, Vec3 position);
LUSTRINE_WRAPPER_EXPORT int get_num_sand_particles();
LUSTRINE_WRAPPER_EXPORT int get_grid_cell_size();

LUSTRINE_WRAPPER_EXPORT Vec3 get_gravity();
LUSTRINE_WRAPPER_EXPORT void set_gravity(Vec3 new_gravity);
LUSTRINE_WRAPPER_EXPORT int add_box(Vec3 position, bool is_dynamic, Vec3 half_dimensions);
LUSTRINE_WRAPPER_EXPORT int add_capsule(Vec3 position, float radius, float height);
LUSTRINE_WRAPPER_EXPORT int add_detector_block(Vec3 position, Vec3 half_dims);
LUSTRINE_WRAPPER_EXPORT int check_collision(int body1, int body2);
LUSTRINE_WRAPPER_EXPORT int do_collide(int body);
LUSTRINE_WRAPPER_EXPORT int do_collide_except_for(int body, int exception_id);
LUSTRINE_WRAPPER_EXPORT void check_collisions(int body, int* indices, int* size);
LUSTRINE_WRAPPER_EXPORT int get_num_bodies();
LUSTRINE_WRAPPER_EXPORT void apply_impulse(int body, Vec3 impulse, Vec3 relative_pos);
LUSTRINE_WRAPPER_EXPORT Vec3 get_position(int body);
LUSTRINE_WRAPPER_EXPORT glm::vec3 get_velocity(int body);
LUSTRINE_WRAPPER_EXPORT void set_velocity(int body, Vec3 velocity);
LUSTRINE_WRAPPER_EXPORT void set_position(int body, Vec3 position);
LUSTRINE_WRAPPER_EXPORT void add_velocity(int body, Vec3 velocity);
LUSTRINE_WRAPPER_EXPORT void set_body_no_rotation(int body);
LUSTRINE_WRAPPER_EXPORT void set_body_frixion(int body, float frixion);
LUSTRINE_WRAPPER_EXPORT float get_body_frixion(int body);
LUSTRINE_WRAPPER_EXPORT void set_body_damping(int body, float linear, float angular);
LUSTRINE_WRAPPER_EXPORT float get_body_damping(int body);
LUSTRINE_WRAPPER_EXPORT void set_player_id(int id);
LUSTRINE_WRAPPER_EXPORT void set_player_box_scale(Vec3 scale);
LUSTRINE_WRAPPER_EXPORT int is_grounded(int id);
LUSTRINE_WRAPPER_EXPORT void set_attract_blow_parameters(float attract_radius, float blow_radius, float attract_coeff, float blow_coeff);

LUSTRINE_WRAPPER_EXPORT int add_particle_source(GridWrapper* pattern, Vec3 direction, float freq, int capacity);
LUSTRINE_WRAPPER_EXPORT int add_particle_sink(Vec3 min_pos, Vec3 max_pos, float frequency);
LUSTRINE_WRAPPER_EXPORT void set_source_state(int index, int state);
LUSTRINE_WRAPPER_EXPORT void set_sink_state(int index, int state);
LUSTRINE_WRAPPER_EXPORT int get_source_spawned(int index);
LUSTRINE_WRAPPER_EXPORT int get_sink_despawned(int index);
LUSTRINE_WRAPPER_EXPORT void set_simulate_function(int index);

LUSTRINE_WRAPPER_EXPORT void set_body_gravity(int id, Vec3 gravity);
LUSTRINE_WRAPPER_EXPORT void set_body_no_collision_response(int id);
LUSTRINE_WRAPPER_EXPORT int collide_with_player(int id);
LUSTRINE_WRAPPER_EXPORT void enable_particles_bounding_boxes();
LUSTRINE_WRAPPER_EXPORT void disable_particles_bounding_boxes();
LUSTRINE_WRAPPER_EXPORT void set_player_particles_bounding_spheres_radius_placement(float radius);
LUSTRINE_WRAPPER_EXPORT int query_cell_num_particles(Vec3 min, Vec3 max, bool include_solid);
LUSTRINE_WRAPPER_EXPORT void test_allocate_1gb();
LUSTRINE_WRAPPER_EXPORT void test_deallocate_1gb();

// ---- B200 additions (not in the reference) ----
// 0 = full host<->device sync per call (default), 1 = resident, 2 = lazy (see Lustrine::B200::HostSync)
LUSTRINE_WRAPPER_EXPORT void b200_set_host_sync(int mode);
LUSTRINE_WRAPPER_EXPORT void b200_set_solver_options(int fluid_iterations, int literal_lambda_index, int exact_math);
LUSTRINE_WRAPPER_EXPORT void b200_set_simulate_function(int index);  // 0 sand, 1 sand_v3, 2 fluid, 3 sand_credits
LUSTRINE_WRAPPER_EXPORT float b200_last_step_ms();
}

}  // namespace Wrapper
}  // namespace Lustrine
