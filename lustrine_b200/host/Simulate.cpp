// Simulate.cpp — the Simulate_fun entry points (reference src/Simulate.cpp).  Each keeps the
// reference's call shape — the host rigid-body step first (src/Simulate.cpp:29,157,328), then the
// particle step — but the particle step is one lgpu_step_* call on the GPU.
#include "lustrine/Simulate.hpp"

#include <cstdlib>
#include <iostream>

#include "DeviceState.hpp"

namespace Lustrine {

namespace {
void run(Simulation* s, float dt, int mode) {
    Bullet::simulate_bullet(&s->bullet_physics_simulation, dt, s->ptr_sand_start, s->ptr_sand_end);
    B200::DeviceState* d = static_cast<B200::DeviceState*>(s->gpu);
    if (!d) {
        std::cerr << "lustrine_b200: simulate_* called on a Simulation without device state; there is no CPU fallback" << std::endl;
        std::abort();
    }
    d->step(s, dt, mode);
}
}  // namespace

void simulate_fluid(Simulation* simulation, float dt) { run(simulation, dt, 1); }
void simulate_sand(Simulation* simulation, float dt) { run(simulation, dt, 2); }
void simulate_sand_credits(Simulation* simulation, float dt) { run(simulation, dt, 3); }

// simulate_sand_v1 / _v2 / _v3 (reference src/Simulate.cpp:513-1007) update pairs IN PLACE in index
// order — sequential by construction, marked unstable by the reference's own header
// (src/Simulate.hpp:8-10) — and reach a different fixed point than simulate_sand.  They are not
// built here (out of scope) and are REFUSED, never silently replaced by simulate_sand:
// a caller that selects one gets a diagnostic and the process stops, like every other failure of
// this path (no CPU fallback).
namespace {
[[noreturn]] void refuse(const char* name) {
    std::cout << "lustrine_b200: " << name << " is not implemented on the B200 path (sequential in-place variant of the reference, "
              << "src/Simulate.cpp:513-1007); use simulate_sand, simulate_sand_credits or simulate_fluid" << std::endl;
    std::abort();
}
}  // namespace
void simulate_sand_v1(Simulation*, float) { refuse("simulate_sand_v1"); }
void simulate_sand_v2(Simulation*, float) { refuse("simulate_sand_v2"); }
void simulate_sand_v3(Simulation*, float) { refuse("simulate_sand_v3"); }

}  // namespace Lustrine
