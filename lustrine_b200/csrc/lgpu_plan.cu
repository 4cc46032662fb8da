// lgpu_plan.cu — host-side slab planning of the C ABI (no device code): where the x-slab boundaries go, given the
// per-column particle histogram (SURVEY §8e "rebalanced every R steps from the per-x-plane histogram").  The Python host
// of the tests and of bench.py carries the same logic (lustrine_b200/slabs.py: plan_slabs, guard_columns, slab_capacity);
// tests/test_slabs.py compares the two on random histograms.
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "lgpu_internal.cuh"

extern "C" int lgpu_slab_guard_columns(int margin) { return margin < 0 ? 0 : (margin / 2 > 1 ? margin / 2 : 1); }

extern "C" int lgpu_plan_slabs(const long long* hist, int grid_x, int world, int min_columns, int margin, int* bounds) {
    if (!hist || !bounds || world < 1 || min_columns < 1 || grid_x < world * min_columns) {
        lgpu_set_error("lgpu_plan_slabs: need at least %d cell column(s) per rank (grid_x=%d, world=%d)", min_columns, grid_x, world);
        return LGPU_ERR_ARG;
    }
    std::vector<long long> cum((size_t)grid_x + 1, 0);
    for (int x = 0; x < grid_x; x++) cum[x + 1] = cum[x] + (hist[x] > 0 ? hist[x] : 0);
    const long long total = cum[grid_x];
    int first = 0, last = grid_x;
    if (margin >= 0 && total > 0) {  // cropped to the occupied columns plus the margin, never beyond the grid
        int lo = 0, hi = grid_x - 1;
        while (hist[lo] <= 0) lo++;
        while (hist[hi] <= 0) hi--;
        first = lo - margin > 0 ? lo - margin : 0;
        last = hi + 1 + margin < grid_x ? hi + 1 + margin : grid_x;
        const int shortfall = world * min_columns - (last - first);  // (every rank still owns min_columns columns)
        if (shortfall > 0) {
            first = first - shortfall > 0 ? first - shortfall : 0;
            last = first + world * min_columns < grid_x ? first + world * min_columns : grid_x;
        }
    }
    bounds[0] = first;
    for (int k = 1; k < world; k++) {
        const double target = (double)total * k / world;
        // first boundary x with cum[x] >= target; the closer of x - 1 / x
        int x = 0;
        {
            int a = 0, b = grid_x + 1;  // lower bound over cum[0..grid_x]
            while (a < b) { const int m = (a + b) / 2; if ((double)cum[m] < target) a = m + 1; else b = m; }
            x = a;
        }
        if (x > 0 && fabs((double)cum[x - 1] - target) <= fabs((double)cum[x < grid_x ? x : grid_x] - target)) x -= 1;
        if (x < bounds[k - 1] + min_columns) x = bounds[k - 1] + min_columns;
        if (x > last - (world - k) * min_columns) x = last - (world - k) * min_columns;
        bounds[k] = x;
    }
    bounds[world] = last;
    return LGPU_OK;
}

extern "C" long long lgpu_slab_capacity(const long long* hist, int grid_x, const int* bounds, int world, int ghost_columns, double factor) {
    if (!hist || !bounds || world < 1 || grid_x < 1) return -1;
    if (ghost_columns < 1) ghost_columns = 1;
    long long need = 0;
    for (int k = 0; k < world; k++) {
        // (the first / last slab also take what lies outside the planned columns: out-of-grid and beyond-the-crop particles)
        const int lo = k == 0 ? 0 : bounds[k], hi = k == world - 1 ? grid_x : bounds[k + 1];
        long long owned = 0, ghosts = 0;
        for (int x = lo; x < hi; x++) owned += hist[x] > 0 ? hist[x] : 0;
        for (int x = (bounds[k] - ghost_columns > 0 ? bounds[k] - ghost_columns : 0); x < bounds[k]; x++) ghosts += hist[x] > 0 ? hist[x] : 0;
        for (int x = bounds[k + 1]; x < (bounds[k + 1] + ghost_columns < grid_x ? bounds[k + 1] + ghost_columns : grid_x); x++) ghosts += hist[x] > 0 ? hist[x] : 0;
        if (owned + 2 * ghosts > need) need = owned + 2 * ghosts;
    }
    return (long long)((double)need * factor) + 4096;
}
