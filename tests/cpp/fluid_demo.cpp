// fluid_demo.cpp — a C++ CALLER of the public Lustrine API: the scene of the reference's own fluid experiment
// (experiments/fluid/fluid.cpp:46-118: 40 x 40 x 80 domain, a 3^3 sand block, the level1_physical.vox solids, four
// particle sources and a sink), headless, behind a small C interface for ctypes.  TEST INFRASTRUCTURE.
//
// The SAME source is compiled twice (tests/cpp/Makefile):
//   against the reference's headers and objects            -> oracle/_ref/libfluid_demo_ref.so    (-DDEMO_REFERENCE)
//   against lustrine_b200/host/include and liblustrine_b200 -> tests/cpp/_build/libfluid_demo_b200.so
// so the drop-in claim is checked where it matters: a caller written for the reference compiles and runs unchanged.
// Nothing here touches LevekGL; the renderer calls of the experiment are simply left out.
#include <chrono>
#include <vector>

#include "Lustrine.hpp"
#include "Simulate.hpp"

#ifdef DEMO_REFERENCE
// oracle/ref_harness.cpp: the reference's fluid loop with the delta-p output double-buffered on `simulation`
extern "C" void refh_use_jacobi(void* simulation, int iterations, int literal_lambda_index);
#endif

namespace {
struct Demo {
    Lustrine::Simulation simulation;
    Lustrine::SimulationParameters parameters;
    std::vector<Lustrine::Grid> sand_grids, solid_grids, source_grids;
};
}  // namespace

extern "C" {

// solid_cells: the X*Y*Z cell values of the voxel model as init_grid_from_magika_voxel lays them out
// (src/VoxelLoader.cpp:67-91: index = x*Y*Z + y*Z + z, value = palette index, 0 = empty); pass null for no solids.
// fun: 0 simulate_sand (the experiment as shipped), 1 simulate_fluid (literal), 2 the Jacobi fluid step with `iterations`
// solver iterations and lambdas[neighbour] (reference build: the harness's simulate_fluid_jacobi).
__attribute__((visibility("default"))) void* demo_create(const int* solid_cells, int sx, int sy, int sz, int fun, int iterations) {
    using namespace Lustrine;
    Demo* d = new Demo();
    SimulationParameters& parameters = d->parameters;
    parameters.X = 40.0f; parameters.Y = 40.0f; parameters.Z = 80.0f;
    parameters.particleRadius = 0.5f; parameters.particleDiameter = 1.0f;

    d->sand_grids.resize(1);
    init_grid_box(&parameters, &d->sand_grids[0], 3, 3, 3, glm::vec3(0, 0, 0), glm::vec4(1.0, 0.2, 1.0, 1.0), MaterialType::SAND);
    if (solid_cells) {
        d->solid_grids.resize(1);
        Grid& g = d->solid_grids[0];
        g.X = sx; g.Y = sy; g.Z = sz;
        g.type = MaterialType::SOLID;
        g.num_grid_cells = sx * sy * sz;
        g.cells.assign(solid_cells, solid_cells + g.num_grid_cells);
        g.colors = std::vector<glm::vec4>(g.num_grid_cells, glm::vec4(0.5, 0.5, 0.5, 1.0));
        g.has_one_color_per_cell = true;
        g.num_occupied_grid_cells = 0;
        for (int c : g.cells) g.num_occupied_grid_cells += c != 0;
        g.sparse_solid = true; g.dynamic_solid = false;
        g.position = glm::vec3(0, 0, 0);
    }
    init_simulation(&parameters, &d->simulation, d->sand_grids, d->solid_grids);
    // (the reference clears `count` BYTES of its velocity array, src/Lustrine.cpp — SURVEY F12)
    for (int i = 0; i < d->simulation.ptr_sand_end; i++) d->simulation.velocities[i] = glm::vec3(0, 0, 0);

    d->source_grids.resize(4);
    init_grid_box(&parameters, &d->source_grids[0], 4, 4, 1, {25, 30, 25}, glm::vec4(1.0, 0.2, 1.0, 1.0), MaterialType::SAND);
    init_grid_box(&parameters, &d->source_grids[1], 5, 5, 1, {35, 25, 25}, glm::vec4(1.0, 0.2, 1.0, 1.0), MaterialType::SAND);
    init_grid_box(&parameters, &d->source_grids[2], 5, 5, 1, {5, 25, 25}, glm::vec4(1.0, 0.2, 1.0, 1.0), MaterialType::SAND);
    init_grid_box(&parameters, &d->source_grids[3], 1, 4, 15, {1, 35, 45}, glm::vec4(1.0, 0.2, 1.0, 1.0), MaterialType::SAND);
    add_particle_source(&d->simulation, &d->source_grids[0], {0, 0, 1}, 1.0f / 10.f, -1);
    add_particle_source(&d->simulation, &d->source_grids[1], {0, 0, 1}, 1.0f / 10.f, -1);
    add_particle_source(&d->simulation, &d->source_grids[2], {0, 0, 1}, 1.0f / 10.f, -1);
    add_particle_source(&d->simulation, &d->source_grids[3], {1, 0, 0}, 1.0f / 10.f, -1);
    add_particle_sink(&d->simulation, {0, 0, 0}, {40, 5, 40}, 0.0f);

    if (fun == 0) d->simulation.simulate_fun = simulate_sand;
    else if (fun == 1) d->simulation.simulate_fun = simulate_fluid;
    else {
#ifdef DEMO_REFERENCE
        refh_use_jacobi(&d->simulation, iterations, 0);
#else
        d->simulation.simulate_fun = simulate_fluid;
        B200::set_solver_options(&d->simulation, iterations, false, true);
#endif
    }
#ifndef DEMO_REFERENCE
    if (fun == 1) B200::set_solver_options(&d->simulation, 1, true, true);
#endif
    return d;
}

__attribute__((visibility("default"))) void demo_info(void* h, int* out) {
    Demo* d = (Demo*)h;
    out[0] = d->simulation.num_sand_particles; out[1] = d->simulation.num_solid_particles;
    out[2] = d->simulation.ptr_sand_end - d->simulation.ptr_sand_start; out[3] = d->simulation.total_allocated;
}

// `calls` frames of Lustrine::simulate; per frame the live sand count and the sum of all coordinates.  Returns seconds.
__attribute__((visibility("default"))) double demo_run(void* h, int calls, float dt, int* counts, double* checksums) {
    Demo* d = (Demo*)h;
    auto t0 = std::chrono::steady_clock::now();
    for (int k = 0; k < calls; k++) {
        Lustrine::simulate(&d->simulation, dt);
        const int b = d->simulation.ptr_sand_start, e = d->simulation.ptr_sand_end;
        if (counts) counts[k] = e - b;
        if (checksums) {
            double s = 0.0;
            for (int i = b; i < e; i++) s += (double)d->simulation.positions[i].x + (double)d->simulation.positions[i].y + (double)d->simulation.positions[i].z;
            checksums[k] = s;
        }
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

__attribute__((visibility("default"))) int demo_positions(void* h, float* out) {
    Demo* d = (Demo*)h;
    const int b = d->simulation.ptr_sand_start, e = d->simulation.ptr_sand_end;
    for (int i = b; i < e; i++) { out[3 * (i - b)] = d->simulation.positions[i].x; out[3 * (i - b) + 1] = d->simulation.positions[i].y; out[3 * (i - b) + 2] = d->simulation.positions[i].z; }
    return e - b;
}

__attribute__((visibility("default"))) void demo_destroy(void* h) {
    Demo* d = (Demo*)h;
    Lustrine::clean_simulation(&d->simulation);
    delete d;
}

}  // extern "C"
