// lgpu_neighbors.cuh — neighbour enumeration in the reference's list order, and the staged
// neighbour table the solver passes replay.
//
// The reference materialises std::vector<int> lists (src/neighbors/Neighbors.cpp:306-361 for
// sand "v1", :386-448 for fluid "v0") once per step and every solver loop accumulates over them
// in list order.  Here the same lists, in the same order, are produced once per substep by
// k_build_table (lgpu_neighbors.cu) and stored as 16-bit codes; each solver pass then
//   1. stages the block's neighbourhood of the CURRENT predicted positions in shared memory with
//      a handful of 1-D bulk (TMA) copies — the storage is cell-sorted, so the 27-cell
//      neighbourhood of 256 consecutive particles is at most 9 contiguous ranges (BlkDesc), and
//   2. replays its list: one coalesced 8-byte load per four neighbours, one LDS.128 per neighbour.
// Lists that do not fit (more than M entries, a neighbourhood larger than the stage, more than
// 2048 solids in a window) fall back to a stencil re-walk with the frozen build-time predicate
//  — always correct, only slower.
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

// ------------------------------------------------------------------------------------------
// Stencil walks over the global (cell-sorted) storage.  f(j, r): j >= 0 sorted sand slot,
// j < 0: ~(sorted solid slot); r = 3*(dy+1) + (dx+1) is the stencil column the entry was found in.
// ------------------------------------------------------------------------------------------

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
            const int r = 3 * (dy + 1) + (dx + 1);
            int zlo = max(c.z - 1, 0), zhi = min(c.z + 1, g.gZ - 1);
            int base = y * g.gXZ + x * g.gZ;
            for (int z = zlo; z <= zhi; z++) {
                int cc = base + z;
                int b = v.cell_start[cc], e = v.cell_start[cc + 1];
                for (int u = b; u < e; u++)
                    if (within_h(g, xi0, f3(v.x0[u]))) f(u, r);
                if (v.n_solid) {
                    int sb = v.solid_cell_start[cc], se = v.solid_cell_start[cc + 1];
                    for (int k = sb; k < se; k++)
                        if (within_h(g, xi0, f3(v.solid_pos[k]))) f(~k, r);
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
            const int r = 3 * (dy + 1) + (dx + 1);
            int zlo = max(c.z - 1, 0), zhi = min(c.z + 1, g.gZ - 1);
            int base = y * g.gXZ + x * g.gZ;
            int b = v.cell_start[base + zlo], e = min(v.cell_start[base + zhi + 1], i);
            for (int u = b; u < e; u++)
                if (within_h(g, xi0, f3(v.x0[u]))) f(u, r);
        }
    }
    // phase B: per cell, solids then sand with a larger slot
    for (int dy = -1; dy <= 1; dy++) {
        int y = c.y + dy;
        if (y < 0 || y >= g.gY) continue;
        for (int dx = -1; dx <= 1; dx++) {
            int x = c.x + dx;
            if (x < 0 || x >= g.gX) continue;
            const int r = 3 * (dy + 1) + (dx + 1);
            int zlo = max(c.z - 1, 0), zhi = min(c.z + 1, g.gZ - 1);
            int base = y * g.gXZ + x * g.gZ;
            for (int z = zlo; z <= zhi; z++) {
                int cc = base + z;
                if (v.n_solid) {
                    int sb = v.solid_cell_start[cc], se = v.solid_cell_start[cc + 1];
                    for (int k = sb; k < se; k++)
                        if (within_h(g, xi0, f3(v.solid_pos[k]))) f(~k, r);
                }
                int b = max(v.cell_start[cc], i + 1), e = v.cell_start[cc + 1];
                for (int u = b; u < e; u++)
                    if (within_h(g, xi0, f3(v.x0[u]))) f(u, r);
            }
        }
    }
}

template <bool SAND, class F>
__device__ __forceinline__ void walk(const View& v, int i, F3 xi0, F&& f) {
    if (SAND) walk_sand(v, i, xi0, f);
    else walk_fluid(v, i, xi0, f);
}

// ------------------------------------------------------------------------------------------
// Shared-memory stage: mbarrier + 1-D bulk copies (cp.async.bulk, the TMA engine; SASS UBLKCP)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

// Block prologue of every staged kernel.  Warp 0 loads the block descriptor and starts the bulk
// copies of `src` (the array the neighbours are read from); nobody else waits for that: the other
// warps go on to their own loads and meet the copies at stage_wait().  `d` may only be read after
// stage_wait() (the mbarrier's completion publishes warp 0's descriptor stores).
// Must be called by all threads of the block (contains one __syncthreads, before anything is in flight).
__device__ __forceinline__ void stage_begin(const View& v, const float4* __restrict__ src, BlkDesc& d, uint64_t* bar, float4* stage) {
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(bar, 1);
        stage[0] = make_float4(1.0e15f, 1.0e15f, 1.0e15f, 0.0f);  // dummy: farther than any support radius
    }
    __syncthreads();
    if (tid < 32) {
        const int* gsrc = (const int*)&v.blk[blockIdx.x];
        for (int t = tid; t < (int)(sizeof(BlkDesc) / sizeof(int)); t += 32) ((int*)&d)[t] = gsrc[t];
        __syncwarp();
        if (tid == 0) {
            const bool copy = d.mode == 0 && v.n > 0;
            uint32_t bytes = 0;
            if (copy) for (int m = 0; m < d.nr; m++) bytes += (uint32_t)d.len[m] * 16u;
            mbar_expect_tx(bar, bytes);  // arrive (release); the phase completes at once when there is nothing to copy
            if (copy) for (int m = 0; m < d.nr; m++) bulk_g2s(stage + d.s0[m], src + d.g0[m], (uint32_t)d.len[m] * 16u, bar);
        }
    }
}
__device__ __forceinline__ void stage_wait(uint64_t* bar) { mbar_wait(bar, 0); }

// Programmatic dependent launch between consecutive solver passes: a pass lets the next one start
// launching as soon as all of its own blocks are resident (pdl_trigger at the top), and the next pass
// runs its independent preamble — list length, table rows — before it waits for the previous pass's
// memory (pdl_wait), so that launch latency, ramp-up and the table loads overlap the previous tail.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <class... KArgs, class... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// virtual-slot mode: at most 3 ranges, unused ones have s0 = INT_MAX
__device__ __forceinline__ int decode_virtual(const BlkDesc& d, uint32_t code) {
    const int m = ((int)code >= d.s0[1] ? 1 : 0) + ((int)code >= d.s0[2] ? 1 : 0);
    return (int)code + d.g0[m] - d.s0[m];
}

// sorted sand slot (>= 0) or ~solid slot (< 0) of a table code
__device__ __forceinline__ int decode_code(const BlkDesc& d, uint32_t code) {
    if (code & LGPU_SOLID_CODE) return ~(d.sbase[(code >> 11) & 15] + (int)(code & (LGPU_SOLID_WINDOW - 1)));
    if (d.mode == 1) return decode_virtual(d, code);
    for (int m = 0; m < d.nr; m++)
        if ((int)code >= d.s0[m] && (int)code < d.s0[m] + d.len[m]) return d.g0[m] + (int)code - d.s0[m];
    return 0;
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
    return r;
}

// shared address of stage slot `code`: one IMAD (the compiler's own shift/mask/add sequence takes three)
__device__ __forceinline__ uint32_t slot_addr(uint32_t stage_addr, uint32_t code) {
    uint32_t a;
    asm("mad.lo.u32 %0, %1, 16, %2;" : "=r"(a) : "r"(code), "r"(stage_addr));
    return a;
}

// Replays the table row of particle i (cnt entries).  body(pj, code, k): pj = position (and .w payload)
// of the k-th neighbour, read from the stage (sand; from `src` through L1/L2 when the block is in
// virtual-slot mode) or from the sorted solid array.
// PAD = true: the row is processed in whole groups of four.  The padding codes of a SAND table are
// 0 = the far-away dummy (no contact); those of a FLUID table are the particle's own slot, whose
// zero separation makes every term of the branch-free fluid bodies vanish (lgpu_fluid.cu).
// The whole row (MG groups of four codes) is loaded up front — before the caller waits for the
// stage — so that the table traffic overlaps the bulk copies.
template <int MG>
struct TableRow {
    uint2 w[MG];
    __device__ __forceinline__ void load(const View& v, int i, int cnt) {
        const uint2* __restrict__ col = v.nbr16 + i;
        const int ng = (cnt + 3) >> 2;
#pragma unroll
        for (int g = 0; g < MG; g++) {
            w[g] = make_uint2(0u, 0u);
            if (g < ng) w[g] = col[(size_t)g * v.cap];
        }
    }
};

// Row load that does not wait for the list length: the first EARLY groups are fetched
// unconditionally (stale codes beyond the list are never replayed), the rest once cnt is known.
template <int MG, int EARLY>
__device__ __forceinline__ void load_row_early(TableRow<MG>& row, const View& v, int i) {
    const uint2* __restrict__ col = v.nbr16 + i;
#pragma unroll
    for (int g = 0; g < EARLY; g++) row.w[g] = col[(size_t)g * v.cap];
}
template <int MG, int EARLY>
__device__ __forceinline__ void load_row_rest(TableRow<MG>& row, const View& v, int i, int cnt) {
    const uint2* __restrict__ col = v.nbr16 + i;
    const int ng = (cnt + 3) >> 2;
#pragma unroll
    for (int g = EARLY; g < MG; g++) {
        row.w[g] = make_uint2(0u, 0u);
        if (g < ng) row.w[g] = col[(size_t)g * v.cap];
    }
}

// code number k of a row held in registers (k is not a constant; a switch keeps the row in registers,
// an indexed or select-chain formulation makes the compiler move it to local memory)
__device__ __forceinline__ uint32_t row_code_reg(const TableRow<8>& row, int k) {
    uint32_t pair;
    switch (k >> 1) {
        case 0: pair = row.w[0].x; break;  case 1: pair = row.w[0].y; break;
        case 2: pair = row.w[1].x; break;  case 3: pair = row.w[1].y; break;
        case 4: pair = row.w[2].x; break;  case 5: pair = row.w[2].y; break;
        case 6: pair = row.w[3].x; break;  case 7: pair = row.w[3].y; break;
        case 8: pair = row.w[4].x; break;  case 9: pair = row.w[4].y; break;
        case 10: pair = row.w[5].x; break; case 11: pair = row.w[5].y; break;
        case 12: pair = row.w[6].x; break; case 13: pair = row.w[6].y; break;
        case 14: pair = row.w[7].x; break; default: pair = row.w[7].y; break;
    }
    return (k & 1) ? pair >> 16 : pair & 0xffffu;
}

// STAGED = true: the caller has checked d.mode == 0 (the virtual-slot loop is not instantiated)
template <bool SOLIDS, bool PAD, int MG, bool STAGED = false, class Body>
__device__ __forceinline__ void replay_row(const View& v, const BlkDesc& d, const TableRow<MG>& row, uint32_t stage_addr,
                                           const float4* __restrict__ src, int cnt, Body&& body) {
    const int ng = (cnt + 3) >> 2;
    if (STAGED || d.mode == 0) {
#pragma unroll
        for (int g = 0; g < MG; g++) {
            if (g < ng) {
                const uint32_t code[4] = {row.w[g].x & 0xffffu, row.w[g].x >> 16, row.w[g].y & 0xffffu, row.w[g].y >> 16};
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int k = 4 * g + q;
                    if (!PAD && k >= cnt) break;
                    float4 pj;
                    if (SOLIDS && (code[q] & LGPU_SOLID_CODE)) pj = v.solid_pos[d.sbase[(code[q] >> 11) & 15] + (int)(code[q] & (LGPU_SOLID_WINDOW - 1))];
                    else pj = lds128(slot_addr(stage_addr, code[q]));
                    body(pj, code[q], k);
                }
            }
        }
    } else {
#pragma unroll
        for (int g = 0; g < MG; g++) {
            if (g < ng) {
                const uint32_t code[4] = {row.w[g].x & 0xffffu, row.w[g].x >> 16, row.w[g].y & 0xffffu, row.w[g].y >> 16};
                for (int q = 0; q < 4; q++) {
                    const int k = 4 * g + q;
                    if (k >= cnt) break;
                    float4 pj;
                    if (SOLIDS && (code[q] & LGPU_SOLID_CODE)) pj = v.solid_pos[d.sbase[(code[q] >> 11) & 15] + (int)(code[q] & (LGPU_SOLID_WINDOW - 1))];
                    else pj = src[decode_virtual(d, code[q])];
                    body(pj, code[q], k);
                }
            }
        }
    }
}

// generic width (max_neighbors != 32): groups are loaded one ahead
template <bool SOLIDS, bool PAD, class Body>
__device__ __forceinline__ void replay_table(const View& v, const BlkDesc& d, uint32_t stage_addr,
                                             const float4* __restrict__ src, int i, int cnt, Body&& body) {
    const uint2* __restrict__ col = v.nbr16 + i;
    const int ng = (cnt + 3) >> 2;
    uint2 w = ng > 0 ? col[0] : make_uint2(0u, 0u);
    if (d.mode == 0) {
        for (int g = 0; g < ng; g++) {
            uint2 wn = make_uint2(0u, 0u);
            if (g + 1 < ng) wn = col[(size_t)(g + 1) * v.cap];
            uint32_t code[4] = {w.x & 0xffffu, w.x >> 16, w.y & 0xffffu, w.y >> 16};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int k = 4 * g + q;
                if (!PAD && k >= cnt) break;
                float4 pj;
                if (SOLIDS && (code[q] & LGPU_SOLID_CODE)) pj = v.solid_pos[d.sbase[(code[q] >> 11) & 15] + (int)(code[q] & (LGPU_SOLID_WINDOW - 1))];
                else pj = lds128(slot_addr(stage_addr, code[q]));
                body(pj, code[q], k);
            }
            w = wn;
        }
    } else {
        for (int k = 0; k < cnt; k++) {
            if ((k & 3) == 0 && k) w = col[(size_t)(k >> 2) * v.cap];
            const uint32_t pair = (k & 2) ? w.y : w.x;
            const uint32_t code = (k & 1) ? pair >> 16 : pair & 0xffffu;
            float4 pj;
            if (SOLIDS && (code & LGPU_SOLID_CODE)) pj = v.solid_pos[d.sbase[(code >> 11) & 15] + (int)(code & (LGPU_SOLID_WINDOW - 1))];
            else pj = src[decode_virtual(d, code)];
            body(pj, code, k);
        }
    }
}

// One entry point for the solver kernels: full-row preload for the default table width (32),
// the generic loop otherwise.  Waits for the stage after the row loads were issued.
template <bool SOLIDS, bool PAD, class Body>
__device__ __forceinline__ void replay_neighbors(const View& v, const BlkDesc& d, uint64_t* bar, const float4* stage,
                                                 const float4* __restrict__ src, int i, int cnt, Body&& body) {
    const uint32_t stage_addr = smem_u32(stage);
    if (v.M == 32) {
        TableRow<8> row;
        row.load(v, i, cnt);
        stage_wait(bar);
        replay_row<SOLIDS, PAD, 8>(v, d, row, stage_addr, src, cnt, body);
    } else {
        stage_wait(bar);
        replay_table<SOLIDS, PAD>(v, d, stage_addr, src, i, cnt, body);
    }
}
