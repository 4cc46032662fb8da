// Kernels.cpp — host versions of the SPH kernels of the reference API (src/Kernels.cpp:6-67),
// restated in the reference's fp32 operation order.  Not used by the GPU step (see Kernels.hpp).
#include "lustrine/Kernels.hpp"

#include <cmath>

namespace Lustrine {

float cubic_kernel(const Simulation* s, float r) {  // :6-19
    const float q = (r * s->kernelFactor) / s->kernelRadius;
    if (q > 1.0f) return 0.0f;
    if (q <= 0.5f) {
        const float q2 = q * q, q3 = q2 * q;
        return s->cubic_kernel_k * (6.0f * q3 - 6.0f * q2 + 1.0f);
    }
    return s->cubic_kernel_k * (2.0f * std::pow(1.0f - q, 3.0f));
}

float cubic_kernel(const Simulation* s, glm::vec3& r) { return cubic_kernel(s, glm::length(r)); }

glm::vec3 cubic_kernel_grad(const Simulation* s, const glm::vec3& r) {  // :26-41
    const float rl = glm::length(r) * s->kernelFactor;
    const float q = rl / s->kernelRadius;
    if (!(rl > 1.0e-5 && q <= 1.0)) return glm::vec3(0.0f);
    const glm::vec3 grad_q = (1.0f / (rl * s->kernelRadius)) * r;
    if (q <= 0.5f) return (s->cubic_kernel_l * q * (3.0f * q - 2.0f)) * grad_q;
    const float f = 1.0f - q;
    return (s->cubic_kernel_l * (-f * f)) * grad_q;
}

float poly6_kernel(const Simulation* s, float r) {  // :43-51, std::pow(float,int) evaluates in double
    if (!(r <= s->kernelRadius)) return 0.0f;
    const float hf = s->kernelRadius * s->kernelFactor;
    return (float)((315.0f / (64.0f * 3.14f * std::pow((double)hf, 9))) *
                   std::pow(std::pow((double)hf, 2) - std::pow((double)(s->kernelFactor * r), 2), 3));
}

// The reference's vec overload passes r.length(), glm's COMPONENT COUNT (3), not the norm.
float poly6_kernel(const Simulation* s, glm::vec3& r) { return poly6_kernel(s, (float)r.length()); }

glm::vec3 spiky_kernel(const Simulation* s, glm::vec3& r) {  // :57-67
    const float rl = glm::length(r);
    if (!(rl > 0.0 && rl <= s->kernelRadius)) return glm::vec3(0.0f);
    const float hf = s->kernelRadius * s->kernelFactor;
    const float temp = (float)((15.0f / (3.14f * std::pow((double)hf, 6))) * std::pow((double)(hf - (rl * s->kernelFactor)), 2));
    return (r / (rl * s->kernelFactor)) * temp;
}

glm::vec3 spiky_kernel_grad(const Simulation* s, const glm::vec3& r) {
    glm::vec3 copy = r;
    return spiky_kernel(s, copy);
}

}  // namespace Lustrine
