#include "lustrine/Profiling.hpp"

namespace Lustrine {
namespace Profiling {

static long long g_cycles[LUSTRINE_MAX_NUM_MEASUREMENTS];
static double g_seconds[LUSTRINE_MAX_NUM_MEASUREMENTS];
static const double kNominalHz = 2700000000.0;  // the reference's hard-coded clock (src/profiling/Profiling.hpp:51)

void init_profiling() {
    for (int i = 0; i < LUSTRINE_MAX_NUM_MEASUREMENTS; i++) { g_cycles[i] = 0; g_seconds[i] = 0.0; }
}
void record(int index, double seconds) {
    if (index < 0 || index >= LUSTRINE_MAX_NUM_MEASUREMENTS) return;
    g_seconds[index] = seconds;
    g_cycles[index] = (long long)(seconds * kNominalHz);
}
extern "C" long get_num_observation() { return LUSTRINE_MAX_NUM_MEASUREMENTS; }
extern "C" long long get_cycles(int index) { return (index >= 0 && index < LUSTRINE_MAX_NUM_MEASUREMENTS) ? g_cycles[index] : 0; }
extern "C" double get_duration(int index) { return (index >= 0 && index < LUSTRINE_MAX_NUM_MEASUREMENTS) ? g_seconds[index] : 0.0; }

}  // namespace Profiling
}  // namespace Lustrine
