// lgpu_neighbors.cuh — neighbour enumeration in the reference's list order.
//
// The reference materialises std::vector<int> lists (src/neighbors/Neighbors.cpp:306-361 for
// sand "v1", :386-448 for fluid "v0") and the solver loops accumulate over them in list order.
// Here the same ORDER is produced by a per-particle stencil walk over the cell-sorted storage;
// k_build_table stores it in a fixed-width column-major table once per substep and the solver
// passes replay the table (or re-walk the stencil with the frozen build-time predicate when a
// list is longer than the table).
//
// Encoding of an entry: sand neighbour = its sorted slot (>= 0); solid neighbour = ~(sorted
// solid slot) (< 0).
#pragma once
#include "lgpu_internal.cuh"

struct CellCoord { int y, x, z; };
__device__ __forceinline__ CellCoord decode_cell(const Geom& g, int id) {
    // src/neighbors/Neighbors.cpp:311-314
    CellCoord c;
    c.y = id / g.gXZ;
    int rem = id - c.y * g.gXZ;
    c.x = rem / g.gZ;
    c.z = rem - c.x * g.gZ;
    return c;
}

// Fluid order (v0): stencil y-outer, x, z-inner (ascending cell id); in each cell the sand
// particles in ascending reference slot (= ascending sorted slot, the sort is stable) and then
// the solids in ascending upload index (cell push order, src/neighbors/Neighbors.cpp:374-384).
// The particle itself is visited once (the reference's lists contain self).
template <class F>
__device__ __forceinline__ void walk_fluid(const View& v, int i, F3 xi0, F&& f) {
    const Geom& g = v.g;
    CellCoord c = decode_cell(g, v.key[i]);
    for (int dy = -1; dy <= 1; dy++) {
        int y = c.y + dy;
        if (y < 0 || y >= g.gY) continue;
        for (int dx = -1; dx <= 1; dx++) {
            int x = c.x + dx;
            if (x < 0 || x >= g.gX) continue;
            int zlo = max(c.z - 1, 0), zhi = min(c.z + 1, g.gZ - 1);
            int base = y * g.gXZ + x * g.gZ;
            for (int z = zlo; z <= zhi; z++) {
                int cc = base + z;
                int b = v.cell_start[cc], e = v.cell_start[cc + 1];
                for (int u = b; u < e; u++)
                    if (within_h(g, xi0, f3(v.x0[u]))) f(u);
                if (v.n_solid) {
                    int sb = v.solid_cell_start[cc], se = v.solid_cell_start[cc + 1];
                    for (int k = sb; k < se; k++)
                        if (within_h(g, xi0, f3(v.solid_pos[k]))) f(~k);
                }
            }
        }
    }
}

// Sand order (v1): the reference fills neighbors[i] with (A) every sand j < i pushed
// symmetrically while j was processed — ascending j, which is ascending cell id, i.e. stencil
// order — and then (B) its own half-stencil walk: per cell the solids first (static cache copied
// before the sand is pushed, src/neighbors/Neighbors.cpp:293,303) and then the sand j >= i
// (:343-353).  Self entries are dropped here because simulate_sand skips them (src/Simulate.cpp:231).
template <class F>
__device__ __forceinline__ void walk_sand(const View& v, int i, F3 xi0, F&& f) {
    const Geom& g = v.g;
    CellCoord c = decode_cell(g, v.key[i]);
    // phase A: sand with a smaller slot (cells up to and including the own cell)
    for (int dy = -1; dy <= 0; dy++) {
        int y = c.y + dy;
        if (y < 0) continue;
        for (int dx = -1; dx <= 1; dx++) {
            int x = c.x + dx;
            if (x < 0 || x >= g.gX) continue;
            int zlo = max(c.z - 1, 0), zhi = min(c.z + 1, g.gZ - 1);
            int base = y * g.gXZ + x * g.gZ;
            int b = v.cell_start[base + zlo], e = min(v.cell_start[base + zhi + 1], i);
            for (int u = b; u < e; u++)
                if (within_h(g, xi0, f3(v.x0[u]))) f(u);
        }
    }
    // phase B: per cell, solids then sand with a larger slot
    for (int dy = -1; dy <= 1; dy++) {
        int y = c.y + dy;
        if (y < 0 || y >= g.gY) continue;
        for (int dx = -1; dx <= 1; dx++) {
            int x = c.x + dx;
            if (x < 0 || x >= g.gX) continue;
            int zlo = max(c.z - 1, 0), zhi = min(c.z + 1, g.gZ - 1);
            int base = y * g.gXZ + x * g.gZ;
            for (int z = zlo; z <= zhi; z++) {
                int cc = base + z;
                if (v.n_solid) {
                    int sb = v.solid_cell_start[cc], se = v.solid_cell_start[cc + 1];
                    for (int k = sb; k < se; k++)
                        if (within_h(g, xi0, f3(v.solid_pos[k]))) f(~k);
                }
                int b = max(v.cell_start[cc], i + 1), e = v.cell_start[cc + 1];
                for (int u = b; u < e; u++)
                    if (within_h(g, xi0, f3(v.x0[u]))) f(u);
            }
        }
    }
}

template <bool SAND, class F>
__device__ __forceinline__ void walk(const View& v, int i, F3 xi0, F&& f) {
    if (SAND) walk_sand(v, i, xi0, f);
    else walk_fluid(v, i, xi0, f);
}

// Replays the neighbour table, or re-walks the stencil when the list did not fit.
template <bool SAND, class F>
__device__ __forceinline__ void for_each_neighbor(const View& v, int i, F&& f) {
    int cnt = v.nbr_cnt[i];
    if (cnt <= v.M) {
        const int* col = v.nbr + i;
        for (int k = 0; k < cnt; k++) f(col[(size_t)k * v.cap]);
    } else {
        walk<SAND>(v, i, f3(v.x0[i]), f);
    }
}
