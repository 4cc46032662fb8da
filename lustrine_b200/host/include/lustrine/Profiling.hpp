// Profiling.hpp — the reference's profiling side channel (src/profiling/Profiling.hpp:14-77):
// 7 slots, read out through three extern "C" getters.  The reference fills the slots from rdtsc at
// a hard-coded 2.7 GHz (:51); here slot 0 (the whole simulate_fun call, src/Lustrine.cpp:802-804)
// is the host wall time of the call, slot 2 the device time of the GPU substep from CUDA events,
// and cycles are reported at the same nominal 2.7 GHz.
#pragma once

#define LUSTRINE_MAX_NUM_MEASUREMENTS 7

namespace Lustrine {
namespace Profiling {
void init_profiling();
void record(int index, double seconds);
extern "C" __attribute__((visibility("default"))) long get_num_observation();
extern "C" __attribute__((visibility("default"))) long long get_cycles(int index);
extern "C" __attribute__((visibility("default"))) double get_duration(int index);
}  // namespace Profiling
}  // namespace Lustrine
