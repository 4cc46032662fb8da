// Kernels.hpp — SPH smoothing kernels of the reference API (src/Kernels.hpp), host versions.
// They exist so that Simulation::W / gradW keep pointing at real functions; the particle step
// itself evaluates the same kernels on the GPU (lustrine_b200/csrc/lgpu_internal.cuh) and only
// looks at WHICH function is wired (cubic vs poly6/spiky).
#pragma once

#include "Simulation.hpp"

namespace Lustrine {
float cubic_kernel(const Simulation* simulation, float r);
float cubic_kernel(const Simulation* simulation, glm::vec3& r);
glm::vec3 cubic_kernel_grad(const Simulation* simulation, const glm::vec3& r);
float poly6_kernel(const Simulation* simulation, float r);
float poly6_kernel(const Simulation* simulation, glm::vec3& r);
glm::vec3 spiky_kernel(const Simulation* simulation, glm::vec3& r);
glm::vec3 spiky_kernel_grad(const Simulation* simulation, const glm::vec3& r);  // const-ref adapter for gradW_fun
}  // namespace Lustrine
