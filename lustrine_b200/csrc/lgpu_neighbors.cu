// lgpu_neighbors.cu — builds the per-particle neighbour table once per substep.
// Replaces the list construction of find_neighbors_uniform_grid (src/neighbors/Neighbors.cpp:386-448)
// and find_neighbors_uniform_grid_v1 (:306-361).
#include "lgpu_neighbors.cuh"

// ------------------------------------------------------------------------------------------
// block descriptors: which contiguous ranges of the sorted storage a block of LGPU_TILE particles
// needs staged (one thread per block; a few thousand threads in all)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_block_ranges(View v, int num_blocks, int stage_slots) {
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
// Appends 16-bit codes to a table row: four consecutive codes share one 8-byte group, groups are
// strided by the capacity (nbr16 layout in lgpu_internal.cuh).
struct RowWriter {
    unsigned short* base;  // code 0 of this particle's row
    uint32_t idx;          // offset of the next code
    uint32_t jump;         // distance between two groups of the same particle, in codes, minus 3
    int M, cnt;
    bool bad;
    __device__ __forceinline__ void init(const View& v, int i) {
        base = reinterpret_cast<unsigned short*>(v.nbr16 + i);
        idx = 0;
        jump = (uint32_t)v.cap * 4u - 3u;
        M = v.M; cnt = 0; bad = false;
    }
    __device__ __forceinline__ void emit(uint32_t code) {
        if (cnt < M) {
            base[idx] = (unsigned short)code;
            idx += (cnt & 3) == 3 ? jump : 1u;
        }
        cnt++;
    }
    __device__ __forceinline__ void finish() {  // pad the last group with the dummy code
        if (cnt < M) for (int k = cnt & 3; k != 0 && k < 4; k++) base[idx++] = 0;
    }
};

template <bool SAND>
__global__ void __launch_bounds__(LGPU_TILE) k_build_table(View v) {
    extern __shared__ float4 stage[];
    __shared__ BlkDesc d;
    __shared__ uint64_t bar;
    // per thread and stencil column: candidate range [lo, hi) in stage slots, replaced by the 32-bit
    // hit mask once the column has been tested; lo16 keeps the range start for the emit phase
    __shared__ uint32_t seg[9][LGPU_TILE];
    __shared__ unsigned short lo16[9][LGPU_TILE];
    const int tid = threadIdx.x;
    const int i = blockIdx.x * LGPU_TILE + tid;
    stage_begin(v, v.x0, d, &bar, stage);
    if (i >= v.n) return;
    const Geom& g = v.g;
    if (g.slab && (v.flags[i] & LGPU_FLAG_GHOST)) {  // ghost of a neighbouring slab: read by others, never updated here
        v.nbr_cnt[i] = LGPU_CNT_GHOST;
        return;
    }
    const F3 xi = f3(v.x0[i]);
    const int key = v.key[i];
    RowWriter w;
    w.init(v, i);

    if (d.mode == 2) {
        // neighbourhood beyond the 16-bit code space: count only, the solver passes re-walk the stencil
        int cnt = 0;
        walk<SAND>(v, i, xi, [&](int, int) { cnt++; });
        v.nbr_cnt[i] = cnt | LGPU_CNT_WALK;
        atomicAdd(&v.counters[1], 1ULL);
        return;
    }

    const CellCoord c = decode_cell(g, key);
    const int zlo = max(c.z - 1, 0), zhi = min(c.z + 1, g.gZ - 1);
    bool slow = false;  // solids in the 27 cells, or a column with more than 32 candidates
#pragma unroll
    for (int r = 0; r < 9; r++) {
        const int y = c.y + r / 3 - 1, x = c.x + r % 3 - 1;
        uint32_t s = 0;
        if (y >= 0 && y < g.gY && x >= 0 && x < g.gX) {
            const int base = y * g.gXZ + x * g.gZ;
            const int b = v.cell_start[base + zlo], e = v.cell_start[base + zhi + 1];
            if (e > b) s = (uint32_t)(d.slotbase[r] + b) | ((uint32_t)(d.slotbase[r] + e) << 16);
            if (e - b > 32) slow = true;
            if (v.n_solid && v.solid_cell_start[base + zhi + 1] > v.solid_cell_start[base + zlo]) slow = true;
        }
        seg[r][tid] = s;
        lo16[r][tid] = (unsigned short)(s & 0xffffu);
    }
    stage_wait(d, &bar);

    if (!slow) {
        // No solid in the 27 cells: the reference order is simply ascending sorted slot over the 9
        // columns (fluid: self included; sand: self skipped — SURVEY F7).
        // Phase 1: one flattened loop over the thread's own candidate ranges (lanes with different
        // range lengths do not idle); the hits of a column are collected in a 32-bit mask.
        if (d.mode == 0) {
            const uint32_t stage_addr = smem_u32(stage);
            int r = 0;
            uint32_t s = seg[0][tid];
            uint32_t a = stage_addr + (s & 0xffffu) * 16u, aend = stage_addr + (s >> 16) * 16u, hits = 0, bit = 1;
            while (true) {
                while (a >= aend) {
                    seg[r][tid] = hits;
                    if (++r == 9) goto emit_phase;
                    s = seg[r][tid];
                    a = stage_addr + (s & 0xffffu) * 16u; aend = stage_addr + (s >> 16) * 16u; hits = 0; bit = 1;
                }
                const float4 pj = lds128(a);
                if (within_h(g, xi, f3(pj))) hits |= bit;
                bit += bit;
                a += 16u;
            }
        } else {
            // virtual-slot mode: same codes, candidates read from the global storage
            int r = 0;
            uint32_t s = seg[0][tid];
            uint32_t u = s & 0xffffu, end = s >> 16, u0 = u, hits = 0;
            const float4* vsrc = v.x0 - d.slotbase[0];  // slot -> sorted particle of column r
            while (true) {
                while (u >= end) {
                    seg[r][tid] = hits;
                    if (++r == 9) goto emit_phase;
                    s = seg[r][tid];
                    u = s & 0xffffu; end = s >> 16; u0 = u; hits = 0;
                    vsrc = v.x0 - d.slotbase[r];
                }
                hits |= (within_h(g, xi, f3(vsrc[u])) ? 1u : 0u) << (u - u0);
                u++;
            }
        }
    emit_phase:
        // Phase 2: one flattened loop over the hits.
        if (SAND) seg[4][tid] &= ~(1u << ((uint32_t)(d.slotbase[4] + i) - (uint32_t)lo16[4][tid]));
        {
            int r = 0;
            uint32_t m = seg[0][tid], base = lo16[0][tid];
            while (true) {
                while (m == 0) {
                    if (++r == 9) goto done;
                    m = seg[r][tid]; base = lo16[r][tid];
                }
                const uint32_t t = __ffs(m) - 1;
                m &= m - 1;
                w.emit(base + t);
            }
        }
    } else {
        // solids interleave with the sand per cell (or a column is very dense): walk in the
        // reference's nested order over the global storage
        walk<SAND>(v, i, xi, [&](int j, int r) {
            if (j >= 0) w.emit((uint32_t)(d.slotbase[r] + j));
            else {
                int off = ~j - d.sbase[r];
                if (off >= LGPU_SOLID_WINDOW) { w.bad = true; off = 0; }
                w.emit(LGPU_SOLID_CODE | ((uint32_t)r << 11) | (uint32_t)off);
            }
        });
    }
done:
    w.finish();
    int word = w.cnt;
    if (w.cnt > w.M || w.bad) { word |= LGPU_CNT_WALK; atomicAdd(&v.counters[1], 1ULL); }
    v.nbr_cnt[i] = word;
}

static bool g_attr_done = false;
int lgpu_launch_build_table(lgpu_ctx* c, bool sand_order) {
    if (c->n == 0) return LGPU_OK;  // (slab mode: the solver drivers still run the refresh protocol)
    View v = lgpu_make_view(c);
    const int nb = (c->n + LGPU_TILE - 1) / LGPU_TILE;
    const size_t smem = sizeof(float4) * LGPU_STAGE_SLOTS;
    if (!g_attr_done) {
        CUDA_TRY(cudaFuncSetAttribute(k_build_table<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaFuncSetAttribute(k_build_table<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        g_attr_done = true;
    }
    k_block_ranges<<<(nb + 127) / 128, 128, 0, c->stream>>>(v, nb, c->stage_slots);
    if (sand_order) k_build_table<true><<<nb, LGPU_TILE, smem, c->stream>>>(v);
    else k_build_table<false><<<nb, LGPU_TILE, smem, c->stream>>>(v);
    c->launches += 2;
    CUDA_TRY(cudaGetLastError());
    return LGPU_OK;
}


#define LGPU_PRELOAD(f) do { cudaFuncAttributes a; CUDA_TRY(cudaFuncGetAttributes(&a, f)); } while (0)
int lgpu_preload_neighbors() {
    LGPU_PRELOAD(k_block_ranges); LGPU_PRELOAD(k_build_table<true>); LGPU_PRELOAD(k_build_table<false>);
    return LGPU_OK;
}
