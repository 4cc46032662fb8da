// Simulate.hpp — the solver entry points of the reference (src/Simulate.hpp:12-17), each a
// Simulate_fun that can be assigned to Simulation::simulate_fun.  All of them run the step on
// the GPU; none has a CPU fallback.
#pragma once

#include "Simulation.hpp"

#ifndef LUSTRINE_EXPORT
#define LUSTRINE_EXPORT __attribute__((visibility("default")))
#endif

namespace Lustrine {
LUSTRINE_EXPORT void simulate_fluid(Simulation* simulation, float dt);
LUSTRINE_EXPORT void simulate_sand(Simulation* simulation, float dt);
LUSTRINE_EXPORT void simulate_sand_credits(Simulation* simulation, float dt);
// The reference marks _v1/_v2/_v3 unstable/experimental (src/Simulate.hpp:8-10); they are sequential
// in-place pair updates and are outside the hot-path scope.  The symbols are kept and run the
// stable simulate_sand step.
LUSTRINE_EXPORT void simulate_sand_v1(Simulation* simulation, float dt);
LUSTRINE_EXPORT void simulate_sand_v2(Simulation* simulation, float dt);
LUSTRINE_EXPORT void simulate_sand_v3(Simulation* simulation, float dt);
}  // namespace Lustrine
