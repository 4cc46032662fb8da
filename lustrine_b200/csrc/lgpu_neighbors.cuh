// lgpu_neighbors.cuh — neighbour enumeration in the reference's list order, and the brick-staged
// neighbour table the solver passes replay.
//
// The reference materialises std::vector<int> lists (src/neighbors/Neighbors.cpp:306-361 for
// sand "v1", :386-448 for fluid "v0") once per step and every solver loop accumulates over them
// in list order.  Here the same lists, in the same order, are produced once per substep by
// k_build_table (lgpu_neighbors.cu) and stored as 16-bit codes.  All staged kernels (the build and
// every solver pass) are PERSISTENT: two blocks per SM, each a producer warp plus LGPU_NCW consumer
// warps, working through the substep's list of non-empty bricks (LGPU_BY x LGPU_BX x LGPU_BZ cells):
//   producer   takes the next brick from a device-side cursor, writes its descriptor to shared
//              memory and issues one 1-D bulk copy (cp.async.bulk, the TMA engine) per halo column
//              into the free stage buffer — the storage is cell-sorted with z fastest, so a column
//              of BZ + 2 cells is ONE contiguous range — completing on that buffer's "full" mbarrier;
//   consumers  wait for "full", take 32-particle chunks of the brick from a shared counter, replay
//              each particle's list (one coalesced 8-byte load per four neighbours, one LDS.128 per
//              neighbour) and arrive on the buffer's "empty" mbarrier when the brick is done.
// With two stage buffers the copies of brick n+1 overlap the gather of brick n.
// Lists that do not fit (more than M entries, a neighbourhood larger than the stage) fall back to a
// stencil re-walk with the frozen build-time predicate — always correct, only slower.
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
// mbarrier + 1-D bulk copies (cp.async.bulk, the TMA engine; SASS UBLKCP / SYNCS)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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

// Programmatic dependent launch between consecutive kernels of the substep: a kernel lets its successor start
// launching as soon as all of its own blocks are resident (pdl_trigger at the top); the successor's blocks
// become resident as this kernel's blocks retire and its producer warps wait for this kernel's memory
// (pdl_wait) before they touch anything — launch latency and ramp-up overlap the previous kernel's tail.
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

// ------------------------------------------------------------------------------------------
// Brick pipeline
// ------------------------------------------------------------------------------------------
// Descriptor of the brick held by one stage buffer (shared memory; written by the producer warp before it
// arrives on the buffer's "full" barrier).
struct BrickInfo {
    int work;                       // work index of the brick; -1 = no more work (the consumers leave)
    int mode;                       // BrickDesc::mode
    int n_own;
    int solid_base;                 // first stage slot that holds a solid (codes >= solid_base are solids)
    int cy0, cx0, cz0;              // cell coordinates of the brick's first own cell
    int next_chunk;                 // 32-particle chunks handed out so far
    int own_prefix[LGPU_OWN_COLS + 1];                 // particles of the own runs before run q
    int own_g0[LGPU_OWN_COLS], own_s0[LGPU_OWN_COLS];  // first sorted slot / stage slot of own run q
    int col_g0[LGPU_HCOLS], col_s0[LGPU_HCOLS];        // halo column hc: sorted sand slot j sits in stage slot col_s0 + (j - col_g0)
    int scol_g0[LGPU_HCOLS], scol_s0[LGPU_HCOLS];      // ... sorted solid slot k
    int col_len[LGPU_HCOLS], scol_len[LGPU_HCOLS];     // particles / solids of halo column hc
    int cs[LGPU_HCOLS][LGPU_HB];    // table build only: stage slot at every cell boundary of every halo column
};

struct BrickShared {
    uint64_t full[LGPU_NSTAGE], empty[LGPU_NSTAGE];
    BrickInfo info[LGPU_NSTAGE];
};
#define LGPU_BRICK_SMEM (sizeof(float4) * LGPU_STAGE_SLOTS * LGPU_NSTAGE + sizeof(BrickShared) + 128)

__device__ __forceinline__ int warp_incl_scan_i(int x, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    return x;
}

// work index -> brick id: full bricks were appended from the front of brick_work, sparse ones from the back
__device__ __forceinline__ int brick_of_work(const View& v, int w, int n_full) {
    return w < n_full ? v.brick_work[w] : v.brick_work[v.NB - 1 - (w - n_full)];
}

// Producer warp, one brick: fills `info`, (BUILD) derives the descriptor from the cell offsets and publishes it to
// v.brick_desc[w] for the later passes, or (!BUILD) reads it from there; then starts the bulk copies of `src` (sand
// columns) and of the sorted solids into `stage`, completing on `full`.  All 32 lanes call it.
template <bool BUILD>
__device__ __forceinline__ void brick_produce(const View& v, int w, int n_full, BrickInfo& info, float4* stage, uint64_t* full,
                                              const float4* __restrict__ src) {
    const int lane = threadIdx.x & 31;
    const Geom& g = v.g;
    const int brick = brick_of_work(v, w, n_full);
    const int by = brick / (v.nbX * v.nbZ);
    const int rem = brick - by * (v.nbX * v.nbZ);
    const int bx = rem / v.nbZ, bz = rem - bx * v.nbZ;
    const int cy0 = by * LGPU_BY, cx0 = bx * LGPU_BX, cz0 = bz * LGPU_BZ;
    int len0, len1 = 0, slen0 = 0, slen1 = 0;  // lane: halo column `lane`; lanes 0..3 also column 32 + lane
    int mode = 0, n_slots = 0;
    if (BUILD) {
        // sorted slot at every cell boundary of every halo column (cells cz0-1 .. cz0+BZ, clipped to the grid; the
        // cell offsets are linear in the cell id with z fastest, so boundary gZ of a column is the next column's start)
        for (int idx = lane; idx < LGPU_HCOLS * LGPU_HB; idx += 32) {
            const int hc = idx / LGPU_HB, t = idx - hc * LGPU_HB;
            const int cy = cy0 - 1 + hc / LGPU_HX, cx = cx0 - 1 + hc % LGPU_HX;
            int val = 0;
            if (cy >= 0 && cy < g.gY && cx >= 0 && cx < g.gX) val = v.cell_start[cy * g.gXZ + cx * g.gZ + min(max(cz0 - 1 + t, 0), g.gZ)];
            info.cs[hc][t] = val;
        }
        if (v.n_solid) {
            for (int idx = lane; idx < 2 * LGPU_HCOLS; idx += 32) {
                const int hc = idx >> 1, t = (idx & 1) * (LGPU_HB - 1);
                const int cy = cy0 - 1 + hc / LGPU_HX, cx = cx0 - 1 + hc % LGPU_HX;
                int val = 0;
                if (cy >= 0 && cy < g.gY && cx >= 0 && cx < g.gX) val = v.solid_cell_start[cy * g.gXZ + cx * g.gZ + min(max(cz0 - 1 + t, 0), g.gZ)];
                if (idx & 1) info.scol_s0[hc] = val; else info.scol_g0[hc] = val;  // (scol_s0 = range end, for the moment)
            }
        }
        __syncwarp();
        len0 = info.cs[lane][LGPU_HB - 1] - info.cs[lane][0];
        info.col_g0[lane] = info.cs[lane][0];
        if (lane < LGPU_HCOLS - 32) { len1 = info.cs[32 + lane][LGPU_HB - 1] - info.cs[32 + lane][0]; info.col_g0[32 + lane] = info.cs[32 + lane][0]; }
        if (v.n_solid) {
            slen0 = info.scol_s0[lane] - info.scol_g0[lane];
            if (lane < LGPU_HCOLS - 32) slen1 = info.scol_s0[32 + lane] - info.scol_g0[32 + lane];
        } else {
            info.scol_g0[lane] = 0;
            if (lane < LGPU_HCOLS - 32) info.scol_g0[32 + lane] = 0;
        }
    } else {
        const BrickDesc& d = v.brick_desc[w];
        len0 = d.len[lane]; info.col_g0[lane] = d.g0[lane];
        slen0 = d.slen[lane]; info.scol_g0[lane] = d.sg0[lane];
        if (lane < LGPU_HCOLS - 32) {
            len1 = d.len[32 + lane]; info.col_g0[32 + lane] = d.g0[32 + lane];
            slen1 = d.slen[32 + lane]; info.scol_g0[32 + lane] = d.sg0[32 + lane];
        }
        mode = d.mode;
    }
    // stage slots: dummies, sand columns 0..35, solid columns 0..35
    const int inc0 = warp_incl_scan_i(len0, lane);
    const int tot0 = __shfl_sync(0xffffffffu, inc0, 31);
    const int inc1 = warp_incl_scan_i(len1, lane);
    const int tot1 = __shfl_sync(0xffffffffu, inc1, 31);
    const int sinc0 = warp_incl_scan_i(slen0, lane);
    const int stot0 = __shfl_sync(0xffffffffu, sinc0, 31);
    const int sinc1 = warp_incl_scan_i(slen1, lane);
    const int stot1 = __shfl_sync(0xffffffffu, sinc1, 31);
    const int s0 = LGPU_DUMMY_SLOTS + inc0 - len0;
    const int s1 = LGPU_DUMMY_SLOTS + tot0 + inc1 - len1;
    const int solid_base = LGPU_DUMMY_SLOTS + tot0 + tot1;
    const int ss0 = solid_base + sinc0 - slen0;
    const int ss1 = solid_base + stot0 + sinc1 - slen1;
    n_slots = solid_base + stot0 + stot1;
    info.col_s0[lane] = s0; info.col_len[lane] = len0;
    info.scol_s0[lane] = ss0; info.scol_len[lane] = slen0;
    if (lane < LGPU_HCOLS - 32) {
        info.col_s0[32 + lane] = s1; info.col_len[32 + lane] = len1;
        info.scol_s0[32 + lane] = ss1; info.scol_len[32 + lane] = slen1;
    }
    __syncwarp();
    // own runs
    int own_len = 0;
    if (BUILD) {
        if (n_slots > v.stage_slots) mode = 2;
        // boundaries -> stage slots
        for (int idx = lane; idx < LGPU_HCOLS * LGPU_HB; idx += 32) {
            const int hc = idx / LGPU_HB, t = idx - hc * LGPU_HB;
            info.cs[hc][t] = info.col_s0[hc] + (info.cs[hc][t] - info.col_g0[hc]);
        }
        __syncwarp();
        if (lane < LGPU_OWN_COLS) {
            const int hc = (lane / LGPU_BX + 1) * LGPU_HX + (lane % LGPU_BX) + 1;
            const int a = info.cs[hc][1], e = info.cs[hc][LGPU_HB - 2];
            own_len = e - a;
            info.own_s0[lane] = a;
            info.own_g0[lane] = info.col_g0[hc] + (a - info.col_s0[hc]);
        }
    } else {
        if (lane < LGPU_OWN_COLS) {
            const BrickDesc& d = v.brick_desc[w];
            const int hc = (lane / LGPU_BX + 1) * LGPU_HX + (lane % LGPU_BX) + 1;
            own_len = d.own_len[lane];
            const int a = d.own_g0[lane];
            info.own_g0[lane] = a;
            info.own_s0[lane] = info.col_s0[hc] + (a - info.col_g0[hc]);
        }
    }
    const int oinc = warp_incl_scan_i(own_len, lane);
    if (lane < LGPU_OWN_COLS) info.own_prefix[lane + 1] = oinc;
    const int n_own = __shfl_sync(0xffffffffu, oinc, LGPU_OWN_COLS - 1);
    if (lane == 0) {
        info.own_prefix[0] = 0;
        info.work = w; info.mode = mode; info.n_own = n_own; info.solid_base = solid_base;
        info.cy0 = cy0; info.cx0 = cx0; info.cz0 = cz0;
        info.next_chunk = 0;
    }
    if (BUILD) {
        BrickDesc& d = v.brick_desc[w];
        d.g0[lane] = info.col_g0[lane]; d.len[lane] = len0; d.sg0[lane] = info.scol_g0[lane]; d.slen[lane] = slen0;
        if (lane < LGPU_HCOLS - 32) {
            d.g0[32 + lane] = info.col_g0[32 + lane]; d.len[32 + lane] = len1;
            d.sg0[32 + lane] = info.scol_g0[32 + lane]; d.slen[32 + lane] = slen1;
        }
        if (lane < LGPU_OWN_COLS) { d.own_g0[lane] = info.own_g0[lane]; d.own_len[lane] = own_len; }
        if (lane == 0) { d.brick = brick; d.mode = mode; d.n_own = n_own; d.n_slots = n_slots; }
    }
    __syncwarp();  // the descriptor stores of all lanes are ordered before lane 0's arrive (release)
    const bool copy = mode == 0;
    if (lane == 0) mbar_expect_tx(full, copy ? 16u * (uint32_t)(n_slots - LGPU_DUMMY_SLOTS) : 0u);
    __syncwarp();
    if (copy) {
        if (len0 > 0) bulk_g2s(stage + s0, src + info.col_g0[lane], 16u * (uint32_t)len0, full);
        if (len1 > 0) bulk_g2s(stage + s1, src + info.col_g0[32 + lane], 16u * (uint32_t)len1, full);
        if (slen0 > 0) bulk_g2s(stage + ss0, v.solid_pos + info.scol_g0[lane], 16u * (uint32_t)slen0, full);
        if (slen1 > 0) bulk_g2s(stage + ss1, v.solid_pos + info.scol_g0[32 + lane], 16u * (uint32_t)slen1, full);
    }
}

// Carves the dynamic shared memory of a brick kernel: stage buffers (16-byte aligned), then barriers + descriptors.
struct BrickSmem {
    float4* stage;     // LGPU_NSTAGE buffers of LGPU_STAGE_SLOTS slots
    BrickShared* sh;
};
__device__ __forceinline__ BrickSmem brick_smem(unsigned char* raw) {
    BrickSmem m;
    uintptr_t a = ((uintptr_t)raw + 127) & ~(uintptr_t)127;
    m.stage = (float4*)a;
    m.sh = (BrickShared*)(a + sizeof(float4) * LGPU_STAGE_SLOTS * LGPU_NSTAGE);
    return m;
}

// The whole persistent loop of a brick kernel.  chunk(info, stage_of_buffer, q, i, slot): called by the lanes of a
// consumer warp that hold a particle of the current 32-particle chunk (no warp collectives inside): sorted slot i,
// own run q, stage slot `slot`.  `cursor` = this pass's work counter (zero at launch).
template <bool BUILD, class Chunk>
__device__ __forceinline__ void brick_loop(const View& v, const float4* __restrict__ src, int* cursor, unsigned char* smem_raw, Chunk&& chunk) {
    const BrickSmem m = brick_smem(smem_raw);
    BrickShared& sh = *m.sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int b = 0; b < LGPU_NSTAGE; b++) { mbar_init(&sh.full[b], 1); mbar_init(&sh.empty[b], LGPU_NCW); }
        mbar_fence_init();
    }
    if (tid < LGPU_NSTAGE * LGPU_DUMMY_SLOTS)  // far-away dummies: farther than any support radius
        m.stage[(tid / LGPU_DUMMY_SLOTS) * LGPU_STAGE_SLOTS + (tid % LGPU_DUMMY_SLOTS)] = make_float4(1.0e15f, 1.0e15f, 1.0e15f, 0.0f);
    __syncthreads();
    pdl_trigger();
    if (warp == LGPU_NCW) {
        // ---- producer ----
        pdl_wait();  // everything this kernel reads was written by its predecessors on the stream
        const int n_full = v.brick_ctl[0], n_work = n_full + v.brick_ctl[1];
        for (int it = 0;; it++) {
            const int b = it % LGPU_NSTAGE;
            if (it >= LGPU_NSTAGE) mbar_wait(&sh.empty[b], ((it / LGPU_NSTAGE) - 1) & 1);
            int w = 0;
            if (lane == 0) w = atomicAdd(cursor, 1);
            w = __shfl_sync(0xffffffffu, w, 0);
            if (w >= n_work) {
                if (lane == 0) { sh.info[b].work = -1; mbar_expect_tx(&sh.full[b], 0u); }
                break;
            }
            brick_produce<BUILD>(v, w, n_full, sh.info[b], m.stage + (size_t)b * LGPU_STAGE_SLOTS, &sh.full[b], src);
        }
    } else {
        // ---- consumers ----
        for (int it = 0;; it++) {
            const int b = it % LGPU_NSTAGE;
            mbar_wait(&sh.full[b], (it / LGPU_NSTAGE) & 1);
            BrickInfo& info = sh.info[b];
            if (info.work < 0) break;
            const int n_own = info.n_own;
            const float4* stage = m.stage + (size_t)b * LGPU_STAGE_SLOTS;
            for (;;) {
                int c = 0;
                if (lane == 0) c = atomicAdd(&info.next_chunk, 1);
                c = __shfl_sync(0xffffffffu, c, 0);
                const int p0 = c * 32;
                if (p0 >= n_own) break;
                int q = 0;  // own run of the chunk's first particle (the same for all lanes), then of the lane's
#pragma unroll
                for (int s = LGPU_OWN_COLS / 2; s > 0; s >>= 1) if (info.own_prefix[q + s] <= p0) q += s;
                const int p = p0 + lane;
                if (p < n_own) {
                    while (info.own_prefix[q + 1] <= p) q++;
                    const int off = p - info.own_prefix[q];
                    chunk(info, stage, q, info.own_g0[q] + off, info.own_s0[q] + off);
                }
                __syncwarp();
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.empty[b]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Table rows
// ------------------------------------------------------------------------------------------
// The whole row (MG groups of four codes) is loaded up front so that the table traffic is in flight before the
// first neighbour is gathered.  PAD: the row is processed in whole groups of four.  The padding codes of a SAND
// table are 0 = a far-away dummy (no contact); those of a FLUID table are the particle's own slot, whose zero
// separation makes every term of the branch-free fluid bodies vanish (lgpu_fluid.cu).
template <int MG>
struct TableRow {
    uint2 w[MG];
    // groups [0, EARLY) unconditionally (stale codes beyond the list are never replayed), the rest once cnt is known
    template <int EARLY> __device__ __forceinline__ void load_early(const View& v, int i) {
        const uint2* col = v.nbr16 + i;
#pragma unroll
        for (int g = 0; g < EARLY; g++) w[g] = col[(size_t)g * v.cap];
    }
    template <int EARLY> __device__ __forceinline__ void load_rest(const View& v, int i, int cnt) {
        const uint2* col = v.nbr16 + i;
        const int ng = (cnt + 3) >> 2;
#pragma unroll
        for (int g = EARLY; g < MG; g++) {
            w[g] = make_uint2(0u, 0u);
            if (g < ng) w[g] = col[(size_t)g * v.cap];
        }
    }
};

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

// replays a register-resident row: body(pj, code, k), pj = stage[code]
template <bool PAD, int MG, class Body>
__device__ __forceinline__ void replay_row(const TableRow<MG>& row, uint32_t stage_addr, int cnt, Body&& body) {
    const int ng = (cnt + 3) >> 2;
#pragma unroll
    for (int g = 0; g < MG; g++) {
        if (g < ng) {
            const uint32_t code[4] = {row.w[g].x & 0xffffu, row.w[g].x >> 16, row.w[g].y & 0xffffu, row.w[g].y >> 16};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int k = 4 * g + q;
                if (!PAD && k >= cnt) break;
                body(lds128(slot_addr(stage_addr, code[q])), code[q], k);
            }
        }
    }
}

// any table width: groups are loaded one ahead
template <bool PAD, class Body>
__device__ __forceinline__ void replay_table(const View& v, uint32_t stage_addr, int i, int cnt, Body&& body) {
    const uint2* col = v.nbr16 + i;
    const int ng = (cnt + 3) >> 2;
    uint2 w = ng > 0 ? col[0] : make_uint2(0u, 0u);
    for (int g = 0; g < ng; g++) {
        uint2 wn = make_uint2(0u, 0u);
        if (g + 1 < ng) wn = col[(size_t)(g + 1) * v.cap];
        const uint32_t code[4] = {w.x & 0xffffu, w.x >> 16, w.y & 0xffffu, w.y >> 16};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int k = 4 * g + q;
            if (!PAD && k >= cnt) break;
            body(lds128(slot_addr(stage_addr, code[q])), code[q], k);
        }
        w = wn;
    }
}

// sorted sand slot (>= 0) or ~(sorted solid slot) (< 0) of a table code of a brick (tests: lgpu_dump)
__device__ __forceinline__ int decode_code(const BrickInfo& info, int code) {
    for (int hc = 0; hc < LGPU_HCOLS; hc++) {
        if (code >= info.col_s0[hc] && code < info.col_s0[hc] + info.col_len[hc]) return info.col_g0[hc] + (code - info.col_s0[hc]);
        if (code >= info.scol_s0[hc] && code < info.scol_s0[hc] + info.scol_len[hc]) return ~(info.scol_g0[hc] + (code - info.scol_s0[hc]));
    }
    return 0;
}
