// DeviceState.hpp — the GPU side of one Lustrine::Simulation: an lgpu context (include/lgpu.h) plus
// the host<->device synchronisation policy.  Internal to liblustrine_b200.so.
#pragma once

#include <vector>

#include "lgpu.h"
#include "lustrine/Lustrine.hpp"

namespace Lustrine {
namespace B200 {

struct DeviceState {
    lgpu_ctx* ctx = nullptr;
    HostSync sync_mode = SYNC_FULL;
    int fluid_iterations = 1;          // the reference's simulate_fluid does one iteration per call
    bool literal_lambda_index = true;  // reference behaviour
    bool exact_math = true;
    bool prev_attract_flag = false;    // the reference keeps this in a function-local static (src/Simulate.cpp:185)
    bool host_pinned = false;          // the live part of the host arrays is page-locked (pin_host)
    void* pinned[4] = {nullptr, nullptr, nullptr, nullptr};
    bool device_flags_valid = true;    // the device holds Simulation::attracted (false after a flag-less upload for a fluid step)
    bool device_matches_host = false;  // the device positions equal the host arrays (set by download, cleared by host-side appends)
    float last_ms = 0.0f;
    int device_sand = 0;               // particles resident on the device
    int capacity = 0;                  // sand slots of the device context (grows on demand, see ensure_capacity)
    int capacity_limit = 0;            // every slot below the solid tail can become sand through the particle sources
    float kernel_radius_scale = 3.1f;

    static DeviceState* create(Simulation* s, float kernel_radius_scale);
    static void destroy(DeviceState* d);

    void make_context(Simulation* s, int capacity_sand);          // (re)creates the lgpu context with that many sand slots
    void ensure_capacity(Simulation* s, int needed);              // grows the context (x1.5) when the sources outgrow it
    void pin_host(Simulation* s);                                // page-locks the first `capacity` sand slots of the host arrays
    void unpin_host();
    void upload(Simulation* s, bool with_flags = true);                                  // host arrays -> device (positions, velocities, attracted)
    void download(Simulation* s, bool with_flags = true);                                // device -> host arrays
    void download_positions_into(float* dst);
    void append_from_host(Simulation* s, int first, int count);  // particle sources
    int remove_in_cells(Simulation* s, const std::vector<int>& cells);
    int cell_count(const int lo[3], const int hi[3], bool include_solid);
    void step(Simulation* s, float dt, int mode /*1 fluid, 2 sand, 3 sand_credits*/);
};

}  // namespace B200
}  // namespace Lustrine
