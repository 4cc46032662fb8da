// oracle/ref_harness.cpp — TEST INFRASTRUCTURE ONLY.
//
// A ctypes-friendly C API around the UNMODIFIED reference (GrapixLeGrand/Lustrine),
// compiled against the reference's own headers and sources where they lie under
// /root/reference (see oracle/Makefile).  Nothing here is shipped or measured as the
// product; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load the resulting oracle/_ref/libref_*.so.
//
// The one permitted deviation from the literal reference is simulate_fluid_jacobi
// below ("O-jac"): the reference's delta-p loop (src/Simulate.cpp:90-113)
// updates positions_star in place while later particles still read it (sequential
// Gauss-Seidel); a data-parallel solver is Jacobi by construction, so fluid positions
// are compared against this double-buffered restatement.  Everything it calls
// (find_neighbors_uniform_grid, W, gradW, s_coor, resolve_collision) is the
// reference's own compiled code.

#include "Lustrine.hpp"
#include "Simulate.hpp"
#include "neighbors/Neighbors.hpp"
#include "neighbors/Sorting.hpp"
#include "neighbors/Utils.hpp"
#include "Kernels.hpp"
#include "BulletPhysics.hpp"

#include <chrono>
#include <cstring>
#include <vector>

namespace Lustrine {
// defined (non-static) in src/Simulate.cpp:7-24 but not declared in any header
float s_coor(const Simulation* simulation, float rl);
float resolve_collision(float value, float min, float max);
}  // namespace Lustrine

namespace {

struct Handle {
    Lustrine::Simulation sim;
    int jacobi_iterations = 1;
    bool literal_lambda_index = true;  // F4: lambdas[j] with j the loop counter
    std::vector<glm::vec3> scratch;
};

Handle* g_current = nullptr;  // simulate_fun has no user pointer; harness is single-sim at a time

// O-jac: src/Simulate.cpp:27-115 with the delta-p output double-buffered, optionally
// repeated K times over the frozen neighbour lists.
void simulate_fluid_jacobi(Lustrine::Simulation* s, float dt) {
    using namespace Lustrine;
    Handle* h = g_current;
    Bullet::simulate_bullet(&s->bullet_physics_simulation, dt, s->ptr_sand_start, s->ptr_sand_end);
    dt = glm::clamp(dt, 0.001f, 0.01f);
    s->time_step = dt;
    const int X = s->domainX, Y = s->domainY, Z = s->domainZ;
    const int b = s->ptr_sand_start, e = s->ptr_sand_end;
    for (int i = b; i < e; i++) {
        s->velocities[i] += s->gravity * s->mass * dt;
        s->positions_star[i] = s->positions[i] + s->velocities[i] * dt;
    }
    find_neighbors_uniform_grid(s);
    h->scratch.resize(e);
    for (int it = 0; it < h->jacobi_iterations; it++) {
        for (int i = b; i < e; i++) {
            const std::vector<int>& nb = s->neighbors[i];
            float rho = 0.0;
            for (size_t j = 0; j < nb.size(); j++) {
                glm::vec3 ij = s->positions_star[i] - s->positions_star[nb[j]];
                rho += s->mass * s->W(s, glm::length(ij));
            }
            rho += s->mass * s->W(s, 0.0);
            float c = (rho / s->rest_density) - 1.0;
            float sum = 0.0;
            glm::vec3 gi = glm::vec3(0.0);
            for (size_t j = 0; j < nb.size(); j++) {
                glm::vec3 t = s->positions_star[i] - s->positions_star[nb[j]];
                glm::vec3 g = -(s->mass / s->rest_density) * s->gradW(s, t);
                sum += glm::dot(g, g);
                gi -= g;
            }
            sum += glm::dot(gi, gi);
            s->lambdas[i] = 0.0;
            if (sum > 0.0) s->lambdas[i] = -c / (sum + s->relaxation_epsilon);
        }
        for (int i = b; i < e; i++) {
            const std::vector<int>& nb = s->neighbors[i];
            glm::vec3 f = glm::vec3(0.0);
            for (size_t j = 0; j < nb.size(); j++) {
                glm::vec3 ij = s->positions_star[i] - s->positions_star[nb[j]];
                float lj = h->literal_lambda_index ? s->lambdas[j] : s->lambdas[nb[j]];
                f += (s->lambdas[i] + lj + s_coor(s, glm::length(ij))) * s->gradW(s, ij);
            }
            f /= s->rest_density;
            glm::vec3 p = s->positions_star[i];
            p += f;
            p.x = resolve_collision(p.x, s->particleRadius, X - s->particleRadius);
            p.y = resolve_collision(p.y, s->particleRadius, Y - s->particleRadius);
            p.z = resolve_collision(p.z, s->particleRadius, Z - s->particleRadius);
            h->scratch[i] = p;
        }
        for (int i = b; i < e; i++) s->positions_star[i] = h->scratch[i];
    }
    for (int i = b; i < e; i++) {
        s->velocities[i] = (s->positions_star[i] - s->positions[i]) / s->time_step;
        s->positions[i] = s->positions_star[i];
    }
}

}  // namespace

extern "C" {

// Creates a simulation through the reference's own init path with n_sand sand particles
// and n_solid solid particles (cell value 2 so that mask_out keeps them).
// Initial coordinates are then overwritten with refh_set_sand / refh_set_solid.
void* refh_create(int X, int Y, int Z, float radius, float diameter, int n_sand, int n_solid,
                  int subdivision, int use_extra, float kernel_radius_scale, int with_credits) {
    using namespace Lustrine;
    Handle* h = new Handle();
    SimulationParameters p;
    p.X = X; p.Y = Y; p.Z = Z;
    p.particleRadius = radius;
    p.particleDiameter = diameter;
    std::vector<Grid> sand, solid;
    if (n_sand > 0) {
        Grid g;
        init_grid_box(&p, &g, n_sand, 1, 1, glm::vec3(0.0f), glm::vec4(1.0f), SAND);
        sand.push_back(g);
    }
    if (n_solid > 0) {
        Grid g;
        init_grid_box(&p, &g, n_solid, 1, 1, glm::vec3(0.0f), glm::vec4(1.0f), SOLID);
        for (auto& c : g.cells) c = 2;
        solid.push_back(g);
    }
    if (use_extra)
        init_simulation_extra_parameters(&p, &h->sim, sand, solid, subdivision, kernel_radius_scale, with_credits != 0);
    else
        init_simulation(&p, &h->sim, sand, solid, subdivision);
    // F12: the reference clears `count` bytes, not floats.
    for (int i = 0; i < h->sim.ptr_solid_ordered_start; i++) h->sim.velocities[i] = glm::vec3(0.0f);
    h->sim.bullet_physics_simulation.player_box_scale = glm::vec3(0.0f);
    g_current = h;
    return h;
}

void refh_destroy(void* hv) {
    Handle* h = (Handle*)hv;
    Lustrine::clean_simulation(&h->sim);
    if (g_current == h) g_current = nullptr;
    delete h;
}

void refh_info(void* hv, int* iv, float* fv) {
    Lustrine::Simulation& s = ((Handle*)hv)->sim;
    iv[0] = s.num_sand_particles; iv[1] = s.num_solid_particles;
    iv[2] = s.ptr_sand_start; iv[3] = s.ptr_sand_end;
    iv[4] = s.ptr_solid_start; iv[5] = s.ptr_solid_end;
    iv[6] = (int)s.total_allocated;
    iv[7] = s.gridX; iv[8] = s.gridY; iv[9] = s.gridZ; iv[10] = s.num_grid_cells;
    iv[11] = s.num_remaining_sand_particles;
    fv[0] = s.kernelRadius; fv[1] = s.cell_size; fv[2] = s.cubic_kernel_k; fv[3] = s.cubic_kernel_l;
    fv[4] = s.domainX; fv[5] = s.domainY; fv[6] = s.domainZ;
    fv[7] = s.particleRadius; fv[8] = s.particleDiameter; fv[9] = s.time_step;
    fv[10] = s.kernelFactor; fv[11] = s.rest_density; fv[12] = s.mass; fv[13] = s.relaxation_epsilon;
    fv[14] = s.bullet_physics_simulation.player_position.x;
    fv[15] = s.bullet_physics_simulation.player_position.y;
    fv[16] = s.bullet_physics_simulation.player_position.z;
}

void refh_set_sand(void* hv, const float* pos, const float* vel, const int* attracted) {
    Lustrine::Simulation& s = ((Handle*)hv)->sim;
    for (int i = s.ptr_sand_start, k = 0; i < s.ptr_sand_end; i++, k++) {
        if (pos) {
            s.positions[i] = glm::vec3(pos[3 * k], pos[3 * k + 1], pos[3 * k + 2]);
            s.positions_star[i] = s.positions[i];
        }
        if (vel) s.velocities[i] = glm::vec3(vel[3 * k], vel[3 * k + 1], vel[3 * k + 2]);
        if (attracted) s.attracted[i] = attracted[k];
    }
}

void refh_set_solid(void* hv, const float* pos) {
    Lustrine::Simulation& s = ((Handle*)hv)->sim;
    for (int i = s.ptr_solid_start, k = 0; i < s.ptr_solid_end; i++, k++) {
        s.positions[i] = glm::vec3(pos[3 * k], pos[3 * k + 1], pos[3 * k + 2]);
        s.positions_star[i] = s.positions[i];
        s.positions_tmp[i] = s.positions[i];
    }
    s.computed_static_particles = false;
    for (auto& c : s.uniform_grid_cells_static_saved) c.clear();
    s.first_iteration = true;
}

void refh_get_sand(void* hv, float* pos, float* pos_star, float* vel, int* attracted) {
    Lustrine::Simulation& s = ((Handle*)hv)->sim;
    const int n = s.ptr_sand_end - s.ptr_sand_start;
    if (pos) memcpy(pos, s.positions + s.ptr_sand_start, sizeof(float) * 3 * n);
    if (pos_star) memcpy(pos_star, s.positions_star + s.ptr_sand_start, sizeof(float) * 3 * n);
    if (vel) memcpy(vel, s.velocities + s.ptr_sand_start, sizeof(float) * 3 * n);
    if (attracted) memcpy(attracted, s.attracted + s.ptr_sand_start, sizeof(int) * n);
}

void refh_get_solid(void* hv, float* pos) {
    Lustrine::Simulation& s = ((Handle*)hv)->sim;
    memcpy(pos, s.positions + s.ptr_solid_start, sizeof(float) * 3 * (s.ptr_solid_end - s.ptr_solid_start));
}

// which: 0 simulate_sand, 1 simulate_fluid (literal), 2 simulate_fluid_jacobi (O-jac), 3 simulate_sand_credits
void refh_set_fun(void* hv, int which, int jacobi_iterations, int literal_lambda_index) {
    Handle* h = (Handle*)hv;
    h->jacobi_iterations = jacobi_iterations;
    h->literal_lambda_index = literal_lambda_index != 0;
    switch (which) {
        case 0: h->sim.simulate_fun = Lustrine::simulate_sand; break;
        case 1: h->sim.simulate_fun = Lustrine::simulate_fluid; break;
        case 2: h->sim.simulate_fun = simulate_fluid_jacobi; break;
        case 3: h->sim.simulate_fun = Lustrine::simulate_sand_credits; break;
    }
}

// Points Simulation::W / gradW at the poly6 / spiky pair of src/Kernels.cpp:43-67 (which 1) or back at the cubic
// spline (0).  spiky_kernel takes a non-const reference, so it needs an adapter to fit gradW_fun.
static glm::vec3 spiky_as_gradW(const Lustrine::Simulation* s, const glm::vec3& r) { glm::vec3 t = r; return Lustrine::spiky_kernel(s, t); }
void refh_set_kernel(void* hv, int which) {
    Lustrine::Simulation& s = ((Handle*)hv)->sim;
    if (which == 1) { s.W = static_cast<Lustrine::W_fun>(Lustrine::poly6_kernel); s.gradW = spiky_as_gradW; }
    else { s.W = static_cast<Lustrine::W_fun>(Lustrine::cubic_kernel); s.gradW = Lustrine::cubic_kernel_grad; }
}

void refh_set_scalars(void* hv, const float* gravity, float rest_density, float mass, float relaxation_epsilon,
                      float s_corr_dq, float s_corr_k, float s_corr_n) {
    Lustrine::Simulation& s = ((Handle*)hv)->sim;
    if (gravity) s.gravity = glm::vec3(gravity[0], gravity[1], gravity[2]);
    s.rest_density = rest_density; s.mass = mass; s.relaxation_epsilon = relaxation_epsilon;
    s.s_corr_dq = s_corr_dq; s.s_corr_k = s_corr_k; s.s_corr_n = s_corr_n;
}

// Places a static Bullet body at `pos` and makes it the "player" so that
// set_particles_box_colliders_positions (src/BulletPhysics.cpp:617) publishes that
// position to the sand predict (src/Simulate.cpp:195-206).
void refh_set_player(void* hv, const float* pos, int attract, int blow, float attract_radius, float blow_radius,
                     float attract_coeff, float blow_coeff) {
    Lustrine::Simulation& s = ((Handle*)hv)->sim;
    static int player_body = -1;
    if (pos) {
        int id = Lustrine::Bullet::add_box(&s.bullet_physics_simulation, glm::vec3(pos[0], pos[1], pos[2]), false);
        s.bullet_physics_simulation.player_id = id;
        player_body = id;
    }
    s.attract_flag = attract != 0;
    s.blow_flag = blow != 0;
    s.attract_radius = attract_radius; s.blow_radius = blow_radius;
    s.attract_coeff = attract_coeff; s.blow_coeff = blow_coeff;
}

// A Simulation owned by somebody else (tests/cpp/fluid_demo.cpp compiled against the reference): its simulate_fun
// becomes the O-jac step with these options.
void refh_use_jacobi(void* simulation, int iterations, int literal_lambda_index) {
    static Handle external;  // (only its options and scratch are used; external.sim stays empty)
    external.jacobi_iterations = iterations;
    external.literal_lambda_index = literal_lambda_index != 0;
    g_current = &external;
    ((Lustrine::Simulation*)simulation)->simulate_fun = simulate_fluid_jacobi;
}

// Runs `steps` calls and returns the wall seconds spent inside them.
double refh_step(void* hv, float dt, int steps, int through_simulate) {
    Handle* h = (Handle*)hv;
    g_current = h;
    auto t0 = std::chrono::steady_clock::now();
    for (int k = 0; k < steps; k++) {
        if (through_simulate) Lustrine::simulate(&h->sim, dt);
        else h->sim.simulate_fun(&h->sim, dt);
    }
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// which: 0 = find_neighbors_uniform_grid (v0, fluid), 1 = find_neighbors_uniform_grid_v1 (sand; permutes storage)
void refh_find_neighbors(void* hv, int which) {
    Lustrine::Simulation& s = ((Handle*)hv)->sim;
    if (which == 0) Lustrine::find_neighbors_uniform_grid(&s);
    else Lustrine::find_neighbors_uniform_grid_v1(&s);
}

void refh_get_lambdas(void* hv, float* out) {
    Lustrine::Simulation& s = ((Handle*)hv)->sim;
    memcpy(out, s.lambdas.data() + s.ptr_sand_start, sizeof(float) * (s.ptr_sand_end - s.ptr_sand_start));
}

long refh_neighbor_counts(void* hv, int* counts) {
    Lustrine::Simulation& s = ((Handle*)hv)->sim;
    long total = 0;
    for (int i = s.ptr_sand_start, k = 0; i < s.ptr_sand_end; i++, k++) {
        if (counts) counts[k] = (int)s.neighbors[i].size();
        total += (long)s.neighbors[i].size();
    }
    return total;
}

// Flat neighbour lists in the reference's own list order (self entries included, F7).
void refh_neighbors(void* hv, int* flat) {
    Lustrine::Simulation& s = ((Handle*)hv)->sim;
    long k = 0;
    for (int i = s.ptr_sand_start; i < s.ptr_sand_end; i++)
        for (int j : s.neighbors[i]) flat[k++] = j;
}

// After find_neighbors_uniform_grid_v1 this holds the cell id of every sorted particle
// (src/neighbors/Neighbors.cpp:301-302).
void refh_sorted_cell_ids(void* hv, int* out) {
    Lustrine::Simulation& s = ((Handle*)hv)->sim;
    memcpy(out, s.counting_sort_arrays->particles_sorted_indices, sizeof(int) * s.num_sand_particles);
}

int refh_cell_id(void* hv, float x, float y, float z) {
    return Lustrine::get_cell_id(&((Handle*)hv)->sim, glm::vec3(x, y, z));
}

void refh_cell_ids(void* hv, const float* pos, int n, int* out) {
    Lustrine::Simulation& s = ((Handle*)hv)->sim;
    for (int i = 0; i < n; i++) out[i] = Lustrine::get_cell_id(&s, glm::vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]));
}

// Sorting::counting_sort, src/neighbors/Sorting.cpp:11-33 (counts must hold num_cells+1 ints)
void refh_counting_sort(int* counts, int* keys, long n, long num_cells, int* sorted) {
    Lustrine::Sorting::counting_sort(counts, keys, (size_t)n, (size_t)num_cells, sorted);
}

// kernel tables: which 0 cubic W, 1 poly6(float), 2 s_coor
void refh_scalar_kernel(void* hv, int which, const float* r, int n, float* out) {
    Lustrine::Simulation& s = ((Handle*)hv)->sim;
    for (int i = 0; i < n; i++) {
        if (which == 0) out[i] = Lustrine::cubic_kernel(&s, r[i]);
        else if (which == 1) out[i] = Lustrine::poly6_kernel(&s, r[i]);
        else out[i] = Lustrine::s_coor(&s, r[i]);
    }
}

// which 0 cubic_kernel_grad, 1 spiky_kernel
void refh_vector_kernel(void* hv, int which, const float* r, int n, float* out) {
    Lustrine::Simulation& s = ((Handle*)hv)->sim;
    for (int i = 0; i < n; i++) {
        glm::vec3 v(r[3 * i], r[3 * i + 1], r[3 * i + 2]);
        glm::vec3 g = which == 0 ? Lustrine::cubic_kernel_grad(&s, v) : Lustrine::spiky_kernel(&s, v);
        out[3 * i] = g.x; out[3 * i + 1] = g.y; out[3 * i + 2] = g.z;
    }
}

int refh_query_cell_num_particles(void* hv, const float* lo, const float* hi, int include_solid) {
    return Lustrine::query_cell_num_particles(&((Handle*)hv)->sim, glm::vec3(lo[0], lo[1], lo[2]),
                                              glm::vec3(hi[0], hi[1], hi[2]), include_solid != 0);
}

int refh_add_sink(void* hv, const float* lo, const float* hi, float frequency) {
    return Lustrine::add_particle_sink(&((Handle*)hv)->sim, glm::vec3(lo[0], lo[1], lo[2]),
                                       glm::vec3(hi[0], hi[1], hi[2]), frequency);
}

int refh_add_source_box(void* hv, int nx, int ny, int nz, const float* origin, const float* direction, float freq, int capacity) {
    Lustrine::Simulation& s = ((Handle*)hv)->sim;
    Lustrine::Grid g;
    Lustrine::init_grid_box(&s.parameters_copy, &g, nx, ny, nz, glm::vec3(origin[0], origin[1], origin[2]), glm::vec4(1.0f), Lustrine::SAND);
    return Lustrine::add_particle_source(&s, &g, glm::vec3(direction[0], direction[1], direction[2]), freq, capacity);
}

}  // extern "C"
