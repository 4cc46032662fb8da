// lgpu_neighbors.cuh — neighbour enumeration in the reference's list order, and the brick-staged
// neighbour table the solver passes replay.
//
// The reference materialises std::vector<int> lists (src/neighbors/Neighbors.cpp:306-361 for
// sand "v1", :386-448 for fluid "v0") once per step and every solver loop accumulates over them
// in list order.  Here the same lists, in the same order, are produced once per substep by
// k_build_table (lgpu_neighbors.cu) and stored as 16-bit codes, one contiguous TABLE BLOCK per brick.
// All staged kernels (the build and every solver pass) run LGPU_CTAS_PER_SM blocks of LGPU_BRICK_THREADS threads per
// SM; block b works through the bricks b, b + gridDim.x, ... of the substep's list of non-empty bricks
// (LGPU_BY x LGPU_BX x LGPU_BZ cells).  Per brick:
//   1. the brick's descriptor arrives in shared memory (one bulk copy, cp.async.bulk = the TMA engine, issued while the
//      block was still working on its previous brick);
//   2. every warp issues the bulk copies of a few halo columns of positions — the storage is cell-sorted with z
//      fastest, so a column of BZ + 2 cells is ONE contiguous range; a warp issues one bulk copy per ~57 cycles
//      whichever lane does it, but the warps of a block issue concurrently (tools/micro/stage_rate.cu) — and, in the
//      solver passes, one warp the copy of the brick's TABLE BLOCK (per own particle: sorted slot, list length, codes);
//      all of them complete on one mbarrier;
//   3. each warp replays the lists of its 32-particle chunks: one conflict-free LDS.64 per four codes, one LDS.128 per
//      neighbour.  The threads read NOTHING from global memory; their only global traffic is the stores of their
//      results.  (Table build: the rows are collected in shared memory and the finished block leaves with one bulk store.)
// While one block of an SM waits for its copies the other blocks of the SM gather: the hardware interleaves the
// blocks' pipelines, no hand-made ring is needed (a persistent producer/consumer ring with 3 slots per SM was built
// and measured slower: a slot is held until its LAST chunk is done, which halves the chunks in flight).
// Lists that do not fit (more than M entries, a neighbourhood larger than a slot) fall back to a
// stencil re-walk with the frozen build-time predicate — always correct, only slower.
#pragma once
#include <stdlib.h>

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
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {  // non-blocking
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok;
    for (;;) {
        // (the hardware suspends the warp for up to the time hint; a waiting warp takes no issue slots)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(addr), "r"(parity), "r"(20000u) : "memory");
        if (ok) break;
    }
}
// shared -> global bulk copy (TMA store) of a finished table block, tracked by the issuing thread's bulk group
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }  // the source may be reused
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }        // the writes are done
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ int ld_acquire_s32(const int* p) {
    int v;
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_s32(int* p, int v) {
    asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 r;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(addr));
    return r;
}
__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
    unsigned short r;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(r) : "r"(addr));
    return (uint32_t)r;
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
    return r;
}
// Packed fp32 pairs (sm_100a FADD2 / FMUL2): both components are rounded to nearest like the scalar instructions,
// so the Exact policy may use them.  (x, y) of a float4 loaded with one LDS.128 already is an aligned register pair.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 unpack2(unsigned long long v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ unsigned long long sub2_rn(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long mul2_rn(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// |xi - pj|^2 in the reference's operation order, every operation separately rounded (src/neighbors/Neighbors.cpp:433-435:
// (dx*dx + dy*dy) + dz*dz), the x and y lanes computed as a pair: 6 instructions instead of 8
__device__ __forceinline__ float dist2_exact(unsigned long long xi_xy, float xi_z, float4 pj) {
    const unsigned long long d = sub2_rn(xi_xy, pack2(pj.x, pj.y));
    const float2 sq = unpack2(mul2_rn(d, d));
    const float dz = __fsub_rn(xi_z, pj.z);
    return __fadd_rn(__fadd_rn(sq.x, sq.y), __fmul_rn(dz, dz));
}

// shared address of stage slot `code`: one IMAD (the compiler's own shift/mask/add sequence takes three)
// shared address of a table CODE: codes are byte offsets into the stage (stage slot x 16 < 65536), so the address is a
// plain add that folds into the load's address operand
__device__ __forceinline__ uint32_t code_addr(uint32_t stage_addr, uint32_t code) { return stage_addr + code; }
static_assert(LGPU_STAGE_SLOTS * 16 <= 65536, "a table code is a 16-bit byte offset into the stage");
__device__ __forceinline__ uint32_t slot_addr(uint32_t stage_addr, uint32_t code) {
    uint32_t a;
    asm("mad.lo.u32 %0, %1, 16, %2;" : "=r"(a) : "r"(code), "r"(stage_addr));
    return a;
}

// ------------------------------------------------------------------------------------------
// Brick pipeline
// ------------------------------------------------------------------------------------------
// One ring slot: the staged neighbourhood and the brick's table block (layout of the block: see BrickDesc).
struct BrickSlot {
    float4 pos[LGPU_STAGE_SLOTS];
    uint2 tab[LGPU_ROW_CAP * (1 + LGPU_MG)];
};

// Shared control block of a block: its brick's descriptor (double-buffered: the next brick's descriptor is fetched
// while this one is worked on; the table build fetches whole records = descriptor + cell boundaries) and barriers.
template <bool BUILD> struct BrickRingEntry { BrickDesc d; };
template <> struct BrickRingEntry<true> { BrickDesc d; unsigned short cs[LGPU_HCOLS][LGPU_HB]; };
template <bool BUILD>
struct BrickShared {
    uint64_t full, dfull[2];
    alignas(16) BrickRingEntry<BUILD> ring[2];
    int work[2];                    // work index of the brick whose descriptor is in ring[e]; -1 = the list is exhausted
    int maxg;                       // table build: groups of the longest row collected so far
};
#define LGPU_BRICK_SMEM_OF(BUILD) (sizeof(BrickSlot) + sizeof(BrickShared<BUILD>) + 128)
#define LGPU_BRICK_SMEM LGPU_BRICK_SMEM_OF(false)
#define LGPU_BRICK_SMEM_BUILD LGPU_BRICK_SMEM_OF(true)
static_assert(LGPU_CTAS_PER_SM * (LGPU_BRICK_SMEM_BUILD + 1024) <= 233472, "the blocks of an SM exceed its 228 KB of shared memory");

__device__ __forceinline__ int warp_incl_scan_i(int x, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    return x;
}

// work index -> record: full bricks were appended from the front of brick_rec, sparse ones from the back
__device__ __forceinline__ int rec_of_work(const View& v, int w, int n_full) {
    return w < n_full ? w : v.rec_cap - 1 - (w - n_full);
}

// Starts the bulk copies of warp `pw`'s share of the brick described by d — `src` (sand columns), the sorted solids
// and, in the solver passes (TABLE), the brick's table block — completing on `full` (whose transaction count thread 0
// sets, brick_stage_bytes).  Called by lane 0 of every warp.
__device__ __forceinline__ uint32_t brick_table_bytes(const BrickDesc& d) { return 8u * (uint32_t)d.n_pad * (uint32_t)(1 + d.maxg); }
template <bool TABLE>
__device__ __forceinline__ uint32_t brick_stage_bytes(const BrickDesc& d) {
    return d.mode == 0 ? 16u * (uint32_t)(d.n_slots - LGPU_DUMMY_SLOTS) + (TABLE ? brick_table_bytes(d) : 0u) : 0u;
}
template <bool TABLE>
__device__ __forceinline__ void brick_stage(const View& v, const BrickDesc& d, BrickSlot& slot, uint64_t* full, const float4* __restrict__ src, int pw) {
    if (d.mode != 0) return;
    if (TABLE && pw == LGPU_BRICK_WARPS - 1) bulk_g2s(slot.tab, v.nbr16 + d.tab_off, brick_table_bytes(d), full);
    for (int hc = pw; hc < LGPU_HCOLS; hc += LGPU_BRICK_WARPS) {
        const int4 c = *reinterpret_cast<const int4*>(&d.col[hc]);
        if (c.z > 0) bulk_g2s(slot.pos + c.y, src + c.x, 16u * (uint32_t)c.z, full);
    }
    if (v.n_solid)
        for (int hc = pw; hc < LGPU_HCOLS; hc += LGPU_BRICK_WARPS) {
            const int4 c = *reinterpret_cast<const int4*>(&d.scol[hc]);
            if (c.z > 0) bulk_g2s(slot.pos + c.y, v.solid_pos + c.x, 16u * (uint32_t)c.z, full);
        }
}

// Carves the dynamic shared memory of a brick kernel: the slot (128-byte aligned), then barriers + descriptors.
template <bool BUILD>
struct BrickSmem {
    BrickSlot* slot;
    BrickShared<BUILD>* sh;
};
template <bool BUILD>
__device__ __forceinline__ BrickSmem<BUILD> brick_smem(unsigned char* raw) {
    BrickSmem<BUILD> m;
    uintptr_t a = ((uintptr_t)raw + 127) & ~(uintptr_t)127;
    m.slot = (BrickSlot*)a;
    m.sh = (BrickShared<BUILD>*)(a + sizeof(BrickSlot));
    return m;
}

// own particle number p of a brick -> own run q
__device__ __forceinline__ int brick_run_of(const BrickDesc& d, int p) {
    int q = 0;
#pragma unroll
    for (int s = LGPU_OWN_COLS / 2; s > 0; s >>= 1) if (d.own_prefix[q + s] <= p) q += s;
    return q;
}

// What a consumer lane knows about its particle while it works on a chunk.
struct Chunk {
    const BrickDesc* d;
    const unsigned short (*cs)[LGPU_HB];   // table build: stage slot at every cell boundary of every halo column
    uint32_t stage_addr;   // shared address of the slot's staged positions
    const float4* stage;
    uint32_t row_addr;     // shared address of the lane's first table group (groups are `row_stride` bytes apart)
    uint32_t row_stride;
    uint2* meta;           // the lane's meta word pair in the slot's table block (table build: to be written)
    int p;                 // own particle number in the brick
    int q;                 // own run of the particle (table build and re-walking bricks; -1 otherwise)
    int i;                 // sorted slot
    int slot;              // stage slot
    int word;              // list length | LGPU_CNT_* (solver passes)
};

__device__ __forceinline__ void atom_max_shared(int* p, int v) {
    asm volatile("red.shared.max.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}

// The brick loop of a staged kernel (see the top of this file).
//   BUILD  = table build: chunk(ck) fills the lane's row in the slot's table block, and the block stores the finished
//            table block to global memory with one bulk copy.
//   !BUILD = solver pass: the table block is copied in with the positions.
// chunk is called by the lanes that hold a particle (no warp collectives inside) and returns the number of table
// groups of the row it has written (table build; 0 otherwise).
// The blocks draw their bricks from `cursor` (a device counter of this pass, zeroed with the work list): the list
// holds the full bricks first and the sparse ones last, so the tail of a pass is made of small pieces.  One thread
// draws the NEXT brick and starts the copy of its descriptor while the block works on the current one.
template <bool BUILD, class ChunkFn>
__device__ __forceinline__ void brick_loop(const View& v, const float4* __restrict__ src, int* cursor, unsigned char* smem_raw, ChunkFn&& chunk) {
    const BrickSmem<BUILD> m = brick_smem<BUILD>(smem_raw);
    BrickShared<BUILD>& sh = *m.sh;
    BrickSlot& slot = *m.slot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr uint32_t kRecBytes = (uint32_t)sizeof(BrickRingEntry<BUILD>);
    static_assert(sizeof(BrickRingEntry<true>) == sizeof(BrickRec), "record layout");
    if (tid == 0) {
        mbar_init(&sh.full, 1); mbar_init(&sh.dfull[0], 1); mbar_init(&sh.dfull[1], 1);
        sh.maxg = 0;
        mbar_fence_init();
    }
    if (tid < LGPU_DUMMY_SLOTS) slot.pos[tid] = make_float4(1.0e15f, 1.0e15f, 1.0e15f, 0.0f);  // far-away dummies: farther than any support radius
    __syncthreads();
    pdl_trigger();
    pdl_wait();  // everything this kernel reads was written by its predecessors on the stream
    const int n_full = v.brick_ctl[0], n_work = n_full + v.brick_ctl[1];
    // thread 1 draws work and fetches descriptors (thread 0 and the other lanes 0 are busy issuing the brick's copies)
    auto draw = [&](int e) {
        const int wj = atomicAdd(cursor, 1);
        if (wj < n_work) {
            sh.work[e] = wj;
            mbar_expect_tx(&sh.dfull[e], kRecBytes);
            bulk_g2s(&sh.ring[e], &v.brick_rec[rec_of_work(v, wj, n_full)], kRecBytes, &sh.dfull[e]);
        } else {
            sh.work[e] = -1;
        }
    };
    if (tid == 1) draw(0);
    __syncthreads();
    for (int it = 0;; it++) {
        const int e = it & 1;
        if (sh.work[e] < 0) break;
        mbar_wait(&sh.dfull[e], (uint32_t)(it >> 1) & 1u);
        const BrickDesc& d = sh.ring[e].d;
        if (tid == 0) {
            // (table build: the bulk store of the previous brick's table block must have read the block before anybody
            // writes rows again — and nobody does before this barrier's phase completes, which needs this arrival)
            if (BUILD && it > 0) bulk_wait_read();
            mbar_expect_tx(&sh.full, brick_stage_bytes<!BUILD>(d));
        }
        if (lane == 0) brick_stage<!BUILD>(v, d, slot, &sh.full, src, warp);
        if (tid == 1) draw(e ^ 1);  // (ring[e ^ 1]'s previous user, brick it - 1, is done: everybody passed the barrier below)
        const int n_own = d.n_own;
        Chunk ck;
        ck.d = &d; ck.stage = slot.pos; ck.stage_addr = smem_u32(slot.pos);
        if constexpr (BUILD) ck.cs = sh.ring[e].cs; else ck.cs = nullptr;
        ck.row_stride = 8u * (uint32_t)d.n_pad;
        const bool from_meta = !BUILD && d.mode == 0;
        mbar_wait(&sh.full, (uint32_t)it & 1u);
        int mg = 0;
        for (int p = tid; p < n_own; p += LGPU_BRICK_THREADS) {  // chunks of 32 own particles: warp, warp + LGPU_BRICK_WARPS, ...
            ck.p = p; ck.meta = slot.tab + p; ck.row_addr = smem_u32(slot.tab + d.n_pad + p);
            if (from_meta) {
                const uint2 mw = lds64(smem_u32(slot.tab + p));
                ck.i = (int)mw.x; ck.word = (int)mw.y & ~LGPU_CNT_SLOT_MASK; ck.slot = ((int)mw.y & LGPU_CNT_SLOT_MASK) >> LGPU_CNT_SLOT_SHIFT; ck.q = -1;
            } else {
                const int q = brick_run_of(d, p);
                const int off = p - d.own_prefix[q];
                ck.q = q; ck.i = d.own_g0[q] + off; ck.slot = d.own_s0[q] + off;
                ck.word = BUILD ? 0 : v.nbr_cnt[ck.i];
            }
            mg = max(mg, chunk(ck));
        }
        if (BUILD && d.mode == 0) {
            mg = __reduce_max_sync(0xffffffffu, mg);
            if (lane == 0 && mg > 0) atom_max_shared(&sh.maxg, mg);
            fence_proxy_async();  // this thread's rows are visible to the bulk store of the block
        }
        __syncthreads();  // the brick is done: the staged positions and the descriptor may be reused
        if (BUILD && d.mode == 0 && tid == 0) {
            const int g = sh.maxg;
            v.brick_rec[d.work].d.maxg = g;
            bulk_s2g(v.nbr16 + d.tab_off, slot.tab, 8u * (uint32_t)d.n_pad * (uint32_t)(1 + g));
            bulk_commit();
            sh.maxg = 0;
        }
    }
    if (BUILD && tid == 0) bulk_wait_all();
}

// Replays the lane's table row from the slot's table block: body(pj, code, k), pj = stage[code], for the first
// min(cnt, M) entries of the list.  PAD: the row is processed in whole groups of four.  The padding codes of a SAND table
// are 0 = a far-away dummy (no contact); those of a FLUID table are the particle's own slot, whose zero separation
// makes every term of the branch-free fluid bodies vanish (lgpu_fluid.cu).  One LDS.64 per group (the next group's
// codes are loaded one ahead), one LDS.128 per code.
template <bool PAD, class Body>
__device__ __forceinline__ void replay_row(const Chunk& ck, int cnt, Body&& body) {
    cnt = min(cnt, 4 * LGPU_MG);
    const int ng = (cnt + 3) >> 2;
    if (ng == 0) return;
    uint32_t a = ck.row_addr;
    uint2 w = lds64(a);
    for (int g = 0; g < ng; g++) {
        a += ck.row_stride;
        uint2 wn = w;
        if (g + 1 < ng) wn = lds64(a);
        const uint32_t code[4] = {w.x & 0xffffu, w.x >> 16, w.y & 0xffffu, w.y >> 16};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int k = 4 * g + q;
            if (!PAD && k >= cnt) break;
            body(lds128(code_addr(ck.stage_addr, code[q])), code[q], k);
        }
        w = wn;
    }
}

// ... and the tail of a list longer than the table width from its spill chunk (global memory; a few particles per
// thousand in a collapsing dam break): entries M .. cnt-1.  Called from out-of-line functions only (the hot loops stay small).
template <bool PAD, class Body>
__device__ __forceinline__ void replay_spill(const View& v, int i, uint32_t stage_addr, int cnt, Body&& body) {
    const uint2* sp = v.nbr_spill + (size_t)v.nbr_ovf[i] * (LGPU_SPILL / 4);
    const int ng = (cnt - 4 * LGPU_MG + 3) >> 2;
    for (int g = 0; g < ng; g++) {
        const uint2 w = sp[g];
        const uint32_t code[4] = {w.x & 0xffffu, w.x >> 16, w.y & 0xffffu, w.y >> 16};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int k = 4 * (LGPU_MG + g) + q;
            if (!PAD && k >= cnt) break;
            body(lds128(code_addr(stage_addr, code[q])), code[q], k);
        }
    }
}

// code number k of the lane's row
__device__ __forceinline__ uint32_t row_code(const Chunk& ck, int k) {
    return lds16(ck.row_addr + (uint32_t)(k >> 2) * ck.row_stride + (uint32_t)(k & 3) * 2u);
}

// sorted sand slot (>= 0) or ~(sorted solid slot) (< 0) of a table code of a brick (tests: lgpu_dump)
__device__ __forceinline__ int decode_code(const BrickDesc& d, int code16) {
    const int code = code16 >> 4;  // codes are stage slots x 16
    for (int hc = 0; hc < LGPU_HCOLS; hc++) {
        if (code >= d.col[hc].s0 && code < d.col[hc].s0 + d.col[hc].len) return d.col[hc].g0 + (code - d.col[hc].s0);
        if (code >= d.scol[hc].s0 && code < d.scol[hc].s0 + d.scol[hc].len) return ~(d.scol[hc].g0 + (code - d.scol[hc].s0));
    }
    return 0;
}
