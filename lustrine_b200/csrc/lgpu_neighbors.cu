// lgpu_neighbors.cu — builds the per-particle neighbour table once per substep.
// Replaces the list construction of find_neighbors_uniform_grid (src/neighbors/Neighbors.cpp:386-448)
// and find_neighbors_uniform_grid_v1 (:306-361).
#include <stdlib.h>

#include "lgpu_neighbors.cuh"

// ------------------------------------------------------------------------------------------
// block descriptors: which contiguous ranges of the sorted storage a block of LGPU_TILE particles
// needs staged (one thread per block; a few thousand threads in all)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_block_ranges(View v, int num_blocks, int stage_slots) {
    pdl_trigger();  // the table build may start its own loads (they do not depend on the descriptors)
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= num_blocks) return;
    const Geom& g = v.g;
    const int first = b * LGPU_TILE, last = min(v.n, first + LGPU_TILE) - 1;
    const int kf = v.key[first], kl = v.key[last];
    BlkDesc d;
    int lo[9], hi[9];
#pragma unroll
    for (int r = 0; r < 9; r++) {
        const int off = (r / 3 - 1) * g.gXZ + (r % 3 - 1) * g.gZ;
        int clo = kf + off - 1, chi = kl + off + 1;
        if (chi < 0 || clo > g.C - 1) { lo[r] = hi[r] = 0; d.sbase[r] = 0; continue; }
        clo = max(clo, 0);
        chi = min(chi, g.C - 1);
        lo[r] = v.cell_start[clo];
        hi[r] = v.cell_start[chi + 1];
        d.sbase[r] = v.n_solid ? v.solid_cell_start[clo] : 0;
    }
    // merge overlapping / abutting ranges (they are already ascending in r) and hand out stage slots
    int nr = 0, slots = 1;  // slot 0 = dummy
    int cur_lo = 0, cur_hi = 0;
    bool open = false;
    int member_of[9];
#pragma unroll
    for (int r = 0; r < 9; r++) {
        member_of[r] = -1;
        if (hi[r] <= lo[r]) continue;
        if (open && lo[r] <= cur_hi) {
            cur_hi = max(cur_hi, hi[r]);
        } else {
            if (open) { d.g0[nr] = cur_lo; d.len[nr] = cur_hi - cur_lo; d.s0[nr] = slots; slots += cur_hi - cur_lo; nr++; }
            cur_lo = lo[r]; cur_hi = hi[r]; open = true;
        }
        member_of[r] = nr;
    }
    if (open) { d.g0[nr] = cur_lo; d.len[nr] = cur_hi - cur_lo; d.s0[nr] = slots; slots += cur_hi - cur_lo; nr++; }
    d.mode = 0;
    if (slots > stage_slots) {
        // Too large for the stage (dense neighbour columns).  Virtual slots: one range per stencil
        // row dy spanning its three columns, same 16-bit codes, neighbours read through L1/L2.
        nr = 0; slots = 1;
        for (int t = 0; t < 3; t++) {
            int a = 0x7fffffff, e = 0;
            for (int r = 3 * t; r < 3 * t + 3; r++) {
                member_of[r] = -1;
                if (hi[r] > lo[r]) { a = min(a, lo[r]); e = max(e, hi[r]); member_of[r] = nr; }
            }
            if (e > a) { d.g0[nr] = a; d.len[nr] = e - a; d.s0[nr] = slots; slots += e - a; nr++; }
        }
        d.mode = slots <= LGPU_VIRTUAL_SLOTS ? 1 : 2;
    }
    for (int m = nr; m < 9; m++) { d.g0[m] = 0; d.len[m] = 0; d.s0[m] = 0x7fffffff; }
#pragma unroll
    for (int r = 0; r < 9; r++) d.slotbase[r] = member_of[r] >= 0 ? d.s0[member_of[r]] - d.g0[member_of[r]] : 0;
    d.nr = nr;
    v.blk[b] = d;
}

// ------------------------------------------------------------------------------------------
// table build
// ------------------------------------------------------------------------------------------
// Appends 16-bit codes to a table row.  Four consecutive codes share one 8-byte group and the groups
// are strided by the capacity (nbr16 layout in lgpu_internal.cuh): the codes are packed in a 64-bit
// shift register and every fourth emit stores one group — coalesced across the warp.
struct RowWriter {
    uint2* col;       // next group of this particle's row
    size_t stride;    // capacity
    uint32_t lo, hi;  // shift register: after four emits lo = c0 | c1 << 16, hi = c2 | c3 << 16
    int M, cnt;
    bool bad;
    __device__ __forceinline__ void init(const View& v, int i) {
        col = v.nbr16 + i;
        stride = (size_t)v.cap;
        lo = hi = 0;
        M = v.M; cnt = 0; bad = false;
    }
    __device__ __forceinline__ void shift_in(uint32_t code) {
        lo = __byte_perm(lo, hi, 0x5432);    // (lo >> 16) | (hi << 16)
        hi = __byte_perm(hi, code, 0x5432);  // (hi >> 16) | (code << 16)
    }
    __device__ __forceinline__ void emit(uint32_t code) {
        if (cnt < M) {
            shift_in(code);
            if ((cnt & 3) == 3) { *col = make_uint2(lo, hi); col += stride; }
        }
        cnt++;
    }
    __device__ __forceinline__ void emit_unchecked(uint32_t code) {  // the caller has checked cnt + (codes to come) <= M
        shift_in(code);
        if ((cnt & 3) == 3) { *col = make_uint2(lo, hi); col += stride; }
        cnt++;
    }
    __device__ __forceinline__ void finish(uint32_t pad) {  // completes the last group with the padding code
        if (cnt < M && (cnt & 3)) {
            for (int k = cnt & 3; k < 4; k++) shift_in(pad);
            *col = make_uint2(lo, hi);
        }
    }
};

// Out-of-line paths of the table build: tiles in virtual-slot mode (candidates read from the global
// storage) and particles with solids in their 27 cells or a very dense column (the reference's nested
// order over the global storage).
template <bool SAND>
__device__ __noinline__ int2 build_row_virtual(const View& v, const BlkDesc& d, int i, F3 xi, int key, uint32_t pad) {
    const Geom& g = v.g;
    RowWriter w;
    w.init(v, i);
    const uint32_t self_code = (uint32_t)(d.slotbase[4] + i);
    const CellCoord c = decode_cell(g, key);
    const int zlo = max(c.z - 1, 0), zhi = min(c.z + 1, g.gZ - 1);
    for (int r = 0; r < 9; r++) {
        const int y = c.y + r / 3 - 1, x = c.x + r % 3 - 1;
        if (y < 0 || y >= g.gY || x < 0 || x >= g.gX) continue;
        const int base = y * g.gXZ + x * g.gZ;
        const int b = v.cell_start[base + zlo], e = v.cell_start[base + zhi + 1];
        const uint32_t first = (uint32_t)(d.slotbase[r] + b);
        for (int u = b; u < e; u++) {
            if (!within_h(g, xi, f3(v.x0[u]))) continue;
            const uint32_t code = first + (uint32_t)(u - b);
            if (SAND && code == self_code) continue;
            w.emit(code);
        }
    }
    w.finish(pad);
    return make_int2(w.cnt, w.bad ? 1 : 0);
}
template <bool SAND>
__device__ __noinline__ int2 build_row_walk(const View& v, const BlkDesc& d, int i, F3 xi, uint32_t pad) {
    RowWriter w;
    w.init(v, i);
    walk<SAND>(v, i, xi, [&](int j, int r) {
        if (j >= 0) w.emit((uint32_t)(d.slotbase[r] + j));
        else {
            int off = ~j - d.sbase[r];
            if (off >= LGPU_SOLID_WINDOW) { w.bad = true; off = 0; }
            w.emit(LGPU_SOLID_CODE | ((uint32_t)r << 11) | (uint32_t)off);
        }
    });
    w.finish(pad);
    return make_int2(w.cnt, w.bad ? 1 : 0);
}

template <bool SAND>
__global__ void __launch_bounds__(LGPU_TILE, LGPU_BLOCKS_PER_SM) k_build_table(const __grid_constant__ View v) {
    extern __shared__ float4 stage[];
    __shared__ BlkDesc d;
    __shared__ uint64_t bar;
    const int tid = threadIdx.x;
    const int i = blockIdx.x * LGPU_TILE + tid;
    const int ic = i < v.n ? i : 0;
    const Geom& g = v.g;
    // the thread's own loads first (position, key, the cell offsets of its nine stencil columns):
    // they overlap the descriptor load and the bulk copies
    const float4 x0i = v.x0[ic];
    const int key = v.key[ic];
    const int flags = g.slab ? v.flags[ic] : 0;
    pdl_wait();  // launched as a programmatic dependent of k_block_ranges: the descriptors are needed from here on
    stage_begin(v, v.x0, d, &bar, stage);
    // per stencil column: the candidates are the sorted slots [cb, ce) — the three cells z-1..z+1
    // of a column are contiguous in the sorted storage
    const CellCoord c = decode_cell(g, key);
    const int zlo = max(c.z - 1, 0), zhi = min(c.z + 1, g.gZ - 1);
    int cb[9], ce[9];
    bool slow = false;  // solids in the 27 cells, or a column with more than 32 candidates
#pragma unroll
    for (int r = 0; r < 9; r++) {
        const int y = c.y + r / 3 - 1, x = c.x + r % 3 - 1;
        cb[r] = ce[r] = 0;
        if (y >= 0 && y < g.gY && x >= 0 && x < g.gX) {
            const int base = y * g.gXZ + x * g.gZ;
            cb[r] = v.cell_start[base + zlo]; ce[r] = v.cell_start[base + zhi + 1];
            if (ce[r] - cb[r] > 32) slow = true;
            if (v.n_solid && v.solid_cell_start[base + zhi + 1] > v.solid_cell_start[base + zlo]) slow = true;
        }
    }
    stage_wait(&bar);  // every thread waits: no bulk copy may outlive the block
    if (i >= v.n) return;
    if (flags & LGPU_FLAG_GHOST) {  // ghost of a neighbouring slab: read by others, never updated here
        v.nbr_cnt[i] = LGPU_CNT_GHOST;
        return;
    }
    const F3 xi = f3(x0i);
    if (d.mode == 2) {
        // neighbourhood beyond the 16-bit code space: count only, the solver passes re-walk the stencil
        int cnt = 0;
        walk<SAND>(v, i, xi, [&](int, int) { cnt++; });
        v.nbr_cnt[i] = cnt | LGPU_CNT_WALK;
        atomicAdd(&v.counters[1], 1ULL);
        return;
    }
    const uint32_t self_code = (uint32_t)(d.slotbase[4] + i);
    // padding: sand = the far-away dummy (no contact); fluid = the particle itself (zero separation:
    // every term of the branch-free fluid bodies vanishes)
    const uint32_t pad = SAND ? 0u : self_code;
    int cnt;
    bool bad;
    if (slow) {
        const int2 r = build_row_walk<SAND>(v, d, i, xi, pad);
        cnt = r.x; bad = r.y != 0;
    } else if (d.mode != 0) {
        const int2 r = build_row_virtual<SAND>(v, d, i, xi, key, pad);
        cnt = r.x; bad = r.y != 0;
    } else {
        RowWriter w;
        w.init(v, i);
        // No solid in the 27 cells: the reference order is simply ascending sorted slot over the 9
        // columns (fluid: self included; sand: self skipped — SURVEY F7).  Per column: test the
        // candidates with the Exact predicate into a hit mask, then emit the hits.
        const uint32_t stage_addr = smem_u32(stage);
#pragma unroll
        for (int r = 0; r < 9; r++) {
            const int n = ce[r] - cb[r];
            if (n == 0) continue;
            const uint32_t first = (uint32_t)(d.slotbase[r] + cb[r]);
            // `out` collects one bit per candidate, shifted in from the right: candidate t ends up at
            // bit n-1-t.  The bit is the sign of h2 - r2 (set <=> r2 > h2; the rounded difference has
            // the exact sign and is +0 on equality), so a test costs the 8 separately rounded
            // operations of the reference's predicate plus one FADD and one funnel shift.
            uint32_t out = 0;
            const uint32_t a = slot_addr(stage_addr, first);
#pragma unroll 4
            for (int t = 0; t < n; t++) {
                const float4 pj = lds128(a + 16u * (uint32_t)t);
                const F3 dd = vsub<Exact>(xi, f3(pj));
                out = __funnelshift_l(__float_as_uint(__fsub_rn(g.h2, vdot<Exact>(dd, dd))), out, 1);
            }
            uint32_t m = ~out & (0xffffffffu >> (32 - n));  // hits; candidate t at bit n-1-t
            const uint32_t top = first + (uint32_t)(n - 1);  // code of bit k = top - k
            if (SAND && r == 4) m &= ~(1u << (top - self_code));
            const int hits = __popc(m);
            if (w.cnt + hits > w.M) { w.cnt += hits; w.bad = true; continue; }  // row too long: the solver passes re-walk
            while (m) {  // ascending candidate = descending bit
                const uint32_t k = 31u - (uint32_t)__clz(m);
                m ^= 1u << k;
                w.emit_unchecked(top - k);
            }
        }
        w.finish(pad);
        cnt = w.cnt; bad = w.bad;
    }
    int word = cnt;
    if (cnt > v.M || bad) { word |= LGPU_CNT_WALK; atomicAdd(&v.counters[1], 1ULL); }
    v.nbr_cnt[i] = word;
}

int lgpu_launch_build_table(lgpu_ctx* c, bool sand_order) {
    if (c->n == 0) return LGPU_OK;  // (slab mode: the solver drivers still run the refresh protocol)
    View v = lgpu_make_view(c);
    const int nb = (c->n + LGPU_TILE - 1) / LGPU_TILE;
    const size_t smem = sizeof(float4) * LGPU_STAGE_SLOTS;
    k_block_ranges<<<(nb + 127) / 128, 128, 0, c->stream>>>(v, nb, c->stage_slots);
    static const bool pdl_env = !(getenv("LGPU_PDL") && atoi(getenv("LGPU_PDL")) == 0);
    const bool pdl = pdl_env && !c->phase_timing && !c->use_graph;
    if (sand_order) CUDA_TRY(launch_pdl(k_build_table<true>, nb, LGPU_TILE, smem, c->stream, pdl, v));
    else CUDA_TRY(launch_pdl(k_build_table<false>, nb, LGPU_TILE, smem, c->stream, pdl, v));
    c->launches += 2;
    CUDA_TRY(cudaGetLastError());
    return LGPU_OK;
}


#define LGPU_PRELOAD(f) do { cudaFuncAttributes a; CUDA_TRY(cudaFuncGetAttributes(&a, f)); } while (0)
// Loads every kernel of this file on the CURRENT device and opts the staged ones into their dynamic
// shared memory.  cudaFuncSetAttribute is per device: lgpu_create calls this for every context.
int lgpu_preload_neighbors() {
    const int smem = (int)(sizeof(float4) * LGPU_STAGE_SLOTS);
    CUDA_TRY(cudaFuncSetAttribute(k_build_table<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_TRY(cudaFuncSetAttribute(k_build_table<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    LGPU_PRELOAD(k_block_ranges); LGPU_PRELOAD(k_build_table<true>); LGPU_PRELOAD(k_build_table<false>);
    return LGPU_OK;
}
