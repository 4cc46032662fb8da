// LustrineWrapper.cpp — the C entry points of the reference's DLL (src/LustrineWrapper.cpp) over one
// process-global simulation, forwarding to the C++ API of this library.
#include "lustrine/LustrineWrapper.hpp"

#include <cstdlib>
#include <cstring>
#include <iostream>
#include <new>
#include <string>
#include <vector>

#include "lustrine/RigidBodyHooks.hpp"
#include "lustrine/Simulate.hpp"

namespace Lustrine {
namespace Wrapper {

static Simulation* simulation = nullptr;  // src/LustrineWrapper.hpp:82

static glm::vec3 to_glm(const Vec3& v) { return glm::vec3(v.x, v.y, v.z); }
static glm::vec4 to_glm(const Color& c) { return glm::vec4(c.r, c.g, c.b, c.a); }
static Vec3 from_glm(const glm::vec3& v) { Vec3 r{v.x, v.y, v.z}; return r; }

static void unwrap(const GridWrapper* w, Grid* g) {  // :84-108 (deep copy of caller memory)
    g->cells.assign(w->cells, w->cells + w->num_grid_cells);
    g->has_one_color_per_cell = w->has_one_color_per_cell;
    g->X = w->X; g->Y = w->Y; g->Z = w->Z;
    g->type = (MaterialType)w->type;
    g->color = to_glm(w->color);
    g->num_occupied_grid_cells = w->num_occupied_grid_cells;
    g->num_grid_cells = w->num_grid_cells;
    g->position = to_glm(w->position);
    g->sparse_solid = true;
    g->dynamic_solid = false;
    if (g->has_one_color_per_cell) {
        g->colors.resize(g->num_grid_cells);
        for (int i = 0; i < g->num_grid_cells; i++) g->colors[i] = to_glm(w->colors[i]);
    }
}

static void wrap(const Grid* g, GridWrapper* w) {  // :116-139 (library-owned arrays handed to the caller)
    w->cells = new int[g->num_grid_cells > 0 ? g->num_grid_cells : 1];
    w->colors = nullptr;
    w->has_one_color_per_cell = g->has_one_color_per_cell;
    w->X = g->X; w->Y = g->Y; w->Z = g->Z;
    w->type = g->type;
    w->color = Color{g->color.r, g->color.g, g->color.b, g->color.a};
    w->num_occupied_grid_cells = g->num_occupied_grid_cells;
    w->num_grid_cells = g->num_grid_cells;
    w->position = from_glm(g->position);
    for (int i = 0; i < g->num_grid_cells; i++) w->cells[i] = g->cells[i];
    if (g->has_one_color_per_cell) {
        w->colors = new Color[g->num_grid_cells > 0 ? g->num_grid_cells : 1];
        for (int i = 0; i < g->num_grid_cells; i++) w->colors[i] = Color{g->colors[i].r, g->colors[i].g, g->colors[i].b, g->colors[i].a};
    }
}

static void fill_data(SimulationData* data) {
    data->start_sand_index = simulation->ptr_sand_start; data->end_sand_index = simulation->ptr_sand_end;
    data->start_solid_index = simulation->ptr_solid_start; data->end_solid_index = simulation->ptr_solid_end;
    data->num_sand_particles = simulation->num_sand_particles; data->num_solid_particles = simulation->num_solid_particles;
}

static bool unwrap_all(GridWrapper* sand, int n_sand, GridWrapper* solid, int n_solid, std::vector<Grid>& a, std::vector<Grid>& b) {
    if (n_sand < 0 || n_solid < 0) { std::cout << "Lustrine bad arguments...?" << std::endl; return false; }  // :177-180
    a.resize(n_sand); b.resize(n_solid);
    for (int i = 0; i < n_sand; i++) unwrap(&sand[i], &a[i]);
    for (int i = 0; i < n_solid; i++) unwrap(&solid[i], &b[i]);
    return true;
}

extern "C" {

void init_simulation(const SimulationParameters* parameters, SimulationData* data, GridWrapper* sand_grids, int num_sand_grids,
                     GridWrapper* solid_grids, int num_solid_grids, int subdivision) {
    std::vector<Grid> a, b;
    if (!unwrap_all(sand_grids, num_sand_grids, solid_grids, num_solid_grids, a, b)) return;
    simulation = new Simulation();
    Lustrine::init_simulation(parameters, simulation, a, b, subdivision);
    fill_data(data);
}

void init_simulation_extra_parameters(const SimulationParameters* parameters, SimulationData* data, GridWrapper* sand_grids, int num_sand_grids,
                                      GridWrapper* solid_grids, int num_solid_grids, int subdivision, float kernel_radius_scale, int with_credits) {
    std::vector<Grid> a, b;
    if (!unwrap_all(sand_grids, num_sand_grids, solid_grids, num_solid_grids, a, b)) return;
    simulation = new Simulation();
    Lustrine::init_simulation_extra_parameters(parameters, simulation, a, b, subdivision, kernel_radius_scale, with_credits != 0);
    fill_data(data);
}

void simulate(float dt, bool attract_flag, bool blow_flag) {  // :343-347
    simulation->attract_flag = attract_flag;
    simulation->blow_flag = blow_flag;
    Lustrine::simulate(simulation, dt);
}
void simulate_no_flags(float dt) { Lustrine::simulate(simulation, dt); }

// :377-379 — a copy into caller memory; with lazy host sync this is where the device->host copy happens
void simulation_bind_positions_copy(float* position_ptr) { B200::copy_positions_to(simulation, position_ptr); }

void cleanup_simulation() {
    if (!simulation) return;
    Lustrine::clean_simulation(simulation);
    delete simulation;
    simulation = nullptr;
}

void init_grid_box(const SimulationParameters* parameters, GridWrapper* grid, int X, int Y, int Z, Vec3 position, Color color, int type) {
    Grid g;
    Lustrine::init_grid_box(parameters, &g, X, Y, Z, to_glm(position), to_glm(color), (MaterialType)type);
    wrap(&g, grid);
}

// .vox -> JSON scene description is an offline asset tool (src/VoxelLoader.cpp:123-217), out of scope:
// an empty JSON object is returned so that bindings keep linking.
void read_vox_scene(BindingString* data, const uint8_t*, int64_t) {
    data->length = 2;
    data->data = new char[3];
    std::memcpy(data->data, "{}", 3);
}
void free_string(BindingString* data) { delete[] data->data; data->length = 0; data->data = nullptr; }

void init_grid_magikavoxel(GridWrapper* grid, const char* path, Vec3 position) {
    Grid g;
    init_grid_from_magika_voxel(&g, path, to_glm(position), SOLID);
    wrap(&g, grid);
}

void create_grid(GridWrapper* grid, const wchar_t* path, int type, int pathlen) {  // :537-630
    std::string p;
    for (int i = 0; i < pathlen; i++) p.push_back((char)path[i]);
    Grid g;
    init_grid_from_magika_voxel(&g, p, glm::vec3(0.0f), (MaterialType)type);
    for (int& c : g.cells) c = c != 0;  // create_grid stores booleans; the reference under-allocates here, we do not
    wrap(&g, grid);
}

int get_num_sand_particles() { return simulation->num_sand_particles; }
int get_grid_cell_size() { return simulation->cell_size; }

Vec3 get_gravity() { return from_glm(Bullet::get_gravity(&simulation->bullet_physics_simulation)); }
void set_gravity(Vec3 g) { Bullet::set_gravity(&simulation->bullet_physics_simulation, to_glm(g)); }
int add_box(Vec3 position, bool is_dynamic, Vec3 half) { return Bullet::add_box(&simulation->bullet_physics_simulation, to_glm(position), is_dynamic, to_glm(half)); }
int add_capsule(Vec3 position, float radius, float height) { return Bullet::add_capsule(&simulation->bullet_physics_simulation, to_glm(position), radius, height); }
int add_detector_block(Vec3 position, Vec3 half) {
    int id = Bullet::add_detector_block(&simulation->bullet_physics_simulation, to_glm(position), to_glm(half));
    set_body_gravity(id, Vec3{0.0f, 0.0f, 0.0f});
    return id;
}
int check_collision(int a, int b) { return (int)Bullet::check_collision(&simulation->bullet_physics_simulation, a, b); }
int do_collide(int body) { return (int)Bullet::do_collide(&simulation->bullet_physics_simulation, body); }
int do_collide_except_for(int body, int exception_id) { return (int)Bullet::do_collide_except_for(&simulation->bullet_physics_simulation, body, exception_id); }
void check_collisions(int body, int* indices, int* size) { Bullet::check_collisions(&simulation->bullet_physics_simulation, body, indices, size); }
int get_num_bodies() { return Bullet::get_num_bodies(&simulation->bullet_physics_simulation); }
void apply_impulse(int body, Vec3 impulse, Vec3 rel) { Bullet::apply_impulse(&simulation->bullet_physics_simulation, body, to_glm(impulse), to_glm(rel)); }
Vec3 get_position(int body) { return from_glm(Bullet::get_body_position(&simulation->bullet_physics_simulation, body)); }
glm::vec3 get_velocity(int body) { return Bullet::get_body_velocity(&simulation->bullet_physics_simulation, body); }
void set_velocity(int body, Vec3 v) { Bullet::set_body_velocity(&simulation->bullet_physics_simulation, body, to_glm(v)); }
void set_position(int body, Vec3 p) { Bullet::set_body_position(&simulation->bullet_physics_simulation, body, to_glm(p)); }
void add_velocity(int body, Vec3 v) { Bullet::add_body_velocity(&simulation->bullet_physics_simulation, body, to_glm(v)); }
void set_body_no_rotation(int body) { Bullet::set_body_no_rotation(&simulation->bullet_physics_simulation, body); }
void set_body_frixion(int body, float f) { Bullet::set_body_frixion(&simulation->bullet_physics_simulation, body, f); }
float get_body_frixion(int body) { return Bullet::get_body_frixion(&simulation->bullet_physics_simulation, body); }
void set_body_damping(int body, float linear, float angular) { Bullet::set_body_damping(&simulation->bullet_physics_simulation, body, linear, angular); }
float get_body_damping(int body) { return Bullet::get_body_lin_damping(&simulation->bullet_physics_simulation, body); }
void set_player_id(int id) { simulation->bullet_physics_simulation.player_id = id; }
void set_player_box_scale(Vec3 scale) { simulation->bullet_physics_simulation.player_box_scale = to_glm(scale); }
int is_grounded(int id) { return Bullet::hook_is_grounded(&simulation->bullet_physics_simulation, id); }
void set_attract_blow_parameters(float attract_radius, float blow_radius, float attract_coeff, float blow_coeff) {  // :505-511
    simulation->attract_radius = attract_radius; simulation->blow_radius = blow_radius;
    simulation->attract_coeff = attract_coeff; simulation->blow_coeff = blow_coeff;
}

int add_particle_source(GridWrapper* pattern, Vec3 direction, float freq, int capacity) {
    Grid g;
    unwrap(pattern, &g);
    return Lustrine::add_particle_source(simulation, &g, to_glm(direction), freq, capacity);
}
int add_particle_sink(Vec3 min_pos, Vec3 max_pos, float frequency) { return Lustrine::add_particle_sink(simulation, to_glm(min_pos), to_glm(max_pos), frequency); }
void set_source_state(int index, int state) { Lustrine::set_source_state(simulation, index, state != 0); }
void set_sink_state(int index, int state) { Lustrine::set_sink_state(simulation, index, state != 0); }
int get_source_spawned(int index) { return Lustrine::get_source_spawned(simulation, index); }
int get_sink_despawned(int index) { return Lustrine::get_sink_despawned(simulation, index); }

void set_simulate_function(int index) {  // :666-677
    switch (index) {
        case 0: simulation->simulate_fun = simulate_sand; break;
        case 1:
            // the reference selects simulate_sand_v3 here; that variant is not built on this path
            // (lustrine_b200/host/Simulate.cpp): say so and keep the current function
            std::cout << "lustrine_b200: set_simulate_function(1) = simulate_sand_v3 is not implemented on the B200 path; "
                      << "the simulate function is left unchanged\n";
            break;
        default: std::cout << "Unreckognized input for simulate func " << index << "\n";
    }
}

void set_body_gravity(int id, Vec3 gravity) { Bullet::hook_set_body_gravity(&simulation->bullet_physics_simulation, id, to_glm(gravity)); }
void set_body_no_collision_response(int id) { Bullet::hook_set_body_no_collision_response(&simulation->bullet_physics_simulation, id); }
int collide_with_player(int id) { return check_collision(id, simulation->bullet_physics_simulation.player_id); }
void enable_particles_bounding_boxes() { simulation->bullet_physics_simulation.particles_bounding_box_requested_state = true; }
void disable_particles_bounding_boxes() { simulation->bullet_physics_simulation.particles_bounding_box_requested_state = false; }
void set_player_particles_bounding_spheres_radius_placement(float radius) { simulation->bullet_physics_simulation.player_box_radius = radius; }
int query_cell_num_particles(Vec3 min, Vec3 max, bool include_solid) { return Lustrine::query_cell_num_particles(simulation, to_glm(min), to_glm(max), include_solid); }

static char* g_dummy = nullptr;
void test_allocate_1gb() { g_dummy = static_cast<char*>(::operator new[](1000000000, std::align_val_t{64})); }
void test_deallocate_1gb() { ::operator delete[](g_dummy, std::align_val_t{64}); g_dummy = nullptr; }

void b200_set_host_sync(int mode) { B200::set_host_sync(simulation, (B200::HostSync)mode); }
void b200_set_solver_options(int fluid_iterations, int literal_lambda_index, int exact_math) {
    B200::set_solver_options(simulation, fluid_iterations, literal_lambda_index != 0, exact_math != 0);
}
void b200_set_simulate_function(int index) {
    switch (index) {
        case 0: simulation->simulate_fun = simulate_sand; break;
        case 1: set_simulate_function(1); break;  // refused with a diagnostic, see above
        case 2: simulation->simulate_fun = simulate_fluid; break;
        case 3: simulation->simulate_fun = simulate_sand_credits; break;
        default: break;
    }
}
float b200_last_step_ms() { return B200::last_step_ms(simulation); }

}  // extern "C"

}  // namespace Wrapper
}  // namespace Lustrine
