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
void simulate_sand_v1(Simulation* simulation, float dt) { run(simulation, dt, 2); }
void simulate_sand_v2(Simulation* simulation, float dt) { run(simulation, dt, 2); }
void simulate_sand_v3(Simulation* simulation, float dt) { run(simulation, dt, 2); }

}  // namespace Lustrine
