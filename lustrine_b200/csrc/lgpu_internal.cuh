// lgpu_internal.cuh — shared declarations of liblgpu.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "lgpu.h"

#define LGPU_DEFAULT_MAX_NEIGHBORS 32
#define LGPU_LAMBDA_HEAD 4096  // lambdas[loop counter] table: first slots in reference order
#define LGPU_BLOCK 128
#define LGPU_MAX_MARKS 96

// ---- brick-staged neighbour table (lgpu_neighbors.cuh) ----
// The solver passes and the table build run over BRICKS of LGPU_BY x LGPU_BX cell columns x LGPU_BZ cells (z is the
// fastest cell coordinate, so every column of a brick is ONE contiguous run of the cell-sorted storage).  A brick's
// neighbourhood is the brick plus a one-cell shell: (BY+2)(BX+2) runs of BZ+2 cells, staged in shared memory with one
// 1-D bulk copy (TMA) per run.  2.8 x the brick's own particles are staged (a 384-particle run of the sorted order
// needed 9 x).
#define LGPU_BY 4
#define LGPU_BX 4
#ifndef LGPU_BZ
#define LGPU_BZ 8                                   // (4 x 4 x 8 cells hold ~476 particles of the unit lattice: 15 chunks of 32 for the 16 chunk slots of two rounds of 8 warps)
#endif
#define LGPU_HX (LGPU_BX + 2)
#define LGPU_HCOLS ((LGPU_BY + 2) * (LGPU_BX + 2))  // 36 halo columns
#define LGPU_HB (LGPU_BZ + 3)                       // cell boundaries of a halo column (BZ + 2 cells)
#define LGPU_OWN_COLS (LGPU_BY * LGPU_BX)           // 16
#ifndef LGPU_BRICK_WARPS
#define LGPU_BRICK_WARPS 8                          // warps of a block of the staged kernels (a brick of the unit lattice has 13-19 chunks of 32 particles at BZ = 8)
#endif
#define LGPU_BRICK_THREADS (LGPU_BRICK_WARPS * 32)
#ifndef LGPU_CTAS_PER_SM
#define LGPU_CTAS_PER_SM 3                          // blocks of the staged kernels per SM: while one waits for its copies the others gather
#endif
#ifndef LGPU_STAGE_SLOTS
#define LGPU_STAGE_SLOTS 1700    // float4 slots of the staged neighbourhood of a block (unit lattice: <= 1600 at BZ = 8)
#endif
#ifndef LGPU_ROW_CAP
#define LGPU_ROW_CAP 608         // own particles of a brick whose table block fits the block's shared memory (unit lattice: <= 588 at BZ = 8 but for the 1.6 % of bricks that hold 7 x 7 x 13: those are cut)
#endif
#define LGPU_MG 8                // table groups (of four 16-bit codes) per row: M = 32
#define LGPU_SPILL 32            // codes of a spill chunk: a list of 33 .. 64 entries keeps its tail in one (global memory)
#define LGPU_STABLE_MAX 256      // cells with more members keep arrival order instead of the stable re-rank (k_reorder)
#define LGPU_DUMMY_SLOTS 8       // stage slots 0..7: far-away dummies (padding of sand rows)
#define LGPU_SOLID_WINDOW 2048
#define LGPU_CNT_WALK (1 << 30)   // nbr_cnt flag: the table row is not usable, re-walk the stencil
#define LGPU_CNT_GHOST (1 << 29)  // nbr_cnt flag: ghost particle of a neighbouring slab (not updated here)
#define LGPU_CNT_GHOST_INNER (1 << 28)  // nbr_cnt flag: ghost in the column next to the owned ones of a two-column ghost layer: its list is built and its
                                        // lambda computed here (all its neighbours are present); its position still comes from its owner
#define LGPU_CNT_MASK 0x0fffffff
#define LGPU_CNT_SLOT_SHIFT 16    // meta word of a table block: list length (<= M) | stage slot << 16 | flags
#define LGPU_CNT_SLOT_MASK 0x0fff0000
#define LGPU_MAX_PASSES 40        // work cursors of one substep (table build + 2 K solver passes)

// One non-empty brick of this substep: written by k_brick_desc (one warp per brick), fetched into shared memory by the
// producer warps of the table build and of every solver pass with a bulk copy, two bricks ahead.
// The brick's TABLE BLOCK lives at nbr16 + tab_off (8-byte words): n_pad meta words {sorted slot, list length | stage
// slot << 16 | LGPU_CNT_*} — one per own particle, in the order of the brick's own runs — followed by LGPU_MG groups of
// n_pad code words (four 16-bit stage-slot codes each, list order).  A solver pass brings in the first (1 + maxg) * n_pad
// words with ONE bulk copy.
struct alignas(16) BrickCol { int g0, s0, len, pad; };     // sorted slots [g0, g0 + len) sit in the stage slots from s0 (one LDS.128 per bulk copy)
struct alignas(16) BrickDesc {
    int brick;                                // brick id = (by * nbX + bx) * nbZ + bz
    int mode;                                 // 0 = staged; 2 = neighbourhood or table block larger than a ring slot: its particles re-walk the stencil
    int n_own;                                // particles of the brick itself
    int n_slots;                              // stage slots of the neighbourhood (dummies included)
    int solid_base;                           // first stage slot that holds a solid (codes >= solid_base are solids)
    int maxg;                                 // table groups (of four codes) of the brick's longest row
    int n_pad;                                // n_own rounded up to an even number (16-byte rows of the table block)
    int tab_off;                              // first 8-byte word of the brick's table block in nbr16
    int cy0, cx0, cz0;                        // cell coordinates of the brick's first own cell
    int work;                                 // index of the brick's record in brick_rec
    int own_prefix[LGPU_OWN_COLS + 1];                 // particles of the own runs (inner columns, own cells) before run q
    int own_g0[LGPU_OWN_COLS], own_s0[LGPU_OWN_COLS];  // first sorted slot / stage slot of own run q
    BrickCol col[LGPU_HCOLS];                 // halo column hc, sand
    BrickCol scol[LGPU_HCOLS];                // halo column hc, solids
};
// What k_brick_desc writes per non-empty brick: the descriptor, and for the table build the stage slot at every cell
// boundary of every halo column.  The solver passes fetch only the descriptor.
struct alignas(16) BrickRec {
    BrickDesc d;
    unsigned short cs[LGPU_HCOLS][LGPU_HB];
};

// ---------------------------------------------------------------------------------------
// Device-side view of a context.  Passed BY VALUE to every kernel (lives in the constant
// bank), so scalar parameters are re-read on every launch like the reference re-reads its
// public struct on every call.
// ---------------------------------------------------------------------------------------
struct Geom {
    float domainX, domainY, domainZ;  // float copies of the int domain (src/Lustrine.cpp:105-107)
    int idomX, idomY, idomZ;          // int X = simulation->domainX truncation (src/Simulate.cpp:35-37)
    float radius, diameter;
    float h, h2;                      // kernelRadius and its fp32 square (src/neighbors/Neighbors.cpp:349)
    float cell_size, kernel_factor;
    float cubic_k, cubic_l;
    int gX, gY, gZ, gXZ, C;           // LOCAL grid of this context (== the reference's grid on a single GPU)
    int x_off;                        // global cell x of local column 0 (slabs: x_lo - gw; single GPU: 0)
    int slab;                         // 1 = this context owns the cell columns [x_lo, x_hi) of a larger grid
    int x_lo, x_hi;                   // owned global cell columns (slab mode)
    int gw;                           // ghost columns on either side of the owned ones (slab mode: 1 or 2)
    int gXg;                          // x extent of the reference's (global) grid
    int guard;                        // slab mode: guard columns of an open side of a cropped plan (lgpu_config::slab_guard_columns)
};

// particle flag bits above the reference's `attracted` bits (slab mode only)
#define LGPU_FLAG_DEAD (1 << 30)    // slot no longer holds a particle of this context (emigrated, or a stale ghost)
#define LGPU_FLAG_GHOST (1 << 29)   // ghost copy of a neighbouring slab's boundary particle (read, never updated here)

// One particle in flight between two slabs (migrant or ghost copy): 64 bytes.
struct HaloRec {
    float4 pos, vel, pstar;
    int flags, orig, pad0, pad1;
};

// Peer-visible memory of a slab context.  Everything a neighbour writes lives here, so one IPC handle
// per context is enough.  side 0 = traffic from / to the LEFT neighbour (lower x), 1 = RIGHT.
struct SlabHeader {
    int flag[2];        // sequence number of the last complete message from side s (written by the peer)
    int in_cnt[2][2][2];  // [turn][side][0 = migrants, 1 = ghosts] of the halo message (written by the peer)
    int error;          // set by a wait kernel that timed out
    int pad[21];
};
struct SlabArena {
    SlabHeader* hdr;
    float4* buf[3];       // x0, pa, pb: the solver's position buffers; neighbours store their boundary
                          // particles' new values straight into this context's ghost slots
    float4* rbox[2][2];   // ghost-refresh inboxes [side][turn]: the owners' new values of the ghosts held from side s, in
                          // message order (ghost copies, then the particles that emigrated to that side)
    // halo inboxes [side][turn]: substeps alternate between two sets, so a neighbour that is already one substep ahead
    // never overwrites a message that has not been consumed yet (it cannot get two ahead: it waits for this slab's halo)
    HaloRec* in_mig[2][2];  // migrants received from side s
    HaloRec* in_gho[2][2];  // ghost copies received from side s
};

struct View {
    Geom g;
    int n;        // live sorted particles (owned + ghosts of neighbouring slabs)
    int n_in;     // entries of the unsorted step-boundary storage (== n on a single GPU; slabs: previous
                  // storage incl. dead slots + this step's inbox)
    int n_owned;  // sand particles this context updates (host bookkeeping; kernels test the flag bits)
    // slabs: outboxes filled by the predict kernel ([side]), and the unsorted index of every ghost copy sent
    HaloRec *out_mig[2], *out_gho[2];
    int *gho_src[2], *mig_src[2];  // unsorted index of every ghost copy / migrant sent to side s
    int *out_cnt;   // [0..1] migrants to L/R, [2..3] ghosts to L/R, [4] halo capacity overflows
    int halo_cap, has_nbr[2];
    int *inv;       // slabs: unsorted index -> sorted slot of this step
    int n_solid;
    int cap;      // row stride of the neighbour table
    int M;        // neighbour table width
    // bricks (lgpu_brick.cuh)
    int nbY, nbX, nbZ, NB;    // brick grid
    int stage_slots;          // stage capacity in use (<= LGPU_STAGE_SLOTS; test hook lgpu_set_stage_slots)
    int* brick_ctl;           // [0] full bricks, [1] sparse bricks, [2] table words allocated, [3] spill chunks allocated, [8 + pass] work cursor of each staged kernel
    BrickRec* brick_rec;      // non-empty bricks of this substep: full ones from the front, sparse ones from the back
    int rec_cap;
    // unsorted (pre-reorder) buffers, indexed by the storage slot of the previous step
    float4 *pos_in, *vel_in, *pstar_in;
    int *flags_in, *orig_in;
    // sorted buffers of this step
    float4 *pos, *vel, *x0;  // x0 = predicted positions at grid-build time (frozen)
    float4 *pa, *pb;         // ping-pong predicted positions of the solver iterations
    int *flags, *orig, *perm;
    int *key_in, *rank_in, *tmp_id, *key;
    int4* sort_rec;  // scattered {reference slot, unsorted index, cell key} records of the sort (lives in the unused sorted-velocity buffer)
    int *cell_count, *cell_start;
    // solids, sorted by cell once (src/neighbors/Neighbors.cpp:266-272)
    float4* solid_pos;
    int *solid_orig, *solid_cell_start;
    // neighbour table: one block per brick (see BrickDesc), 16-bit codes in list order, four per uint2.  A code is the
    // stage slot of the neighbour in the particle's brick (sand and solids alike; slots below LGPU_DUMMY_SLOTS are
    // far-away dummies).
    uint2* nbr16;
    int* nbr_cnt;      // by sorted slot: list length | LGPU_CNT_* flags (dumps, re-walking bricks)
    uint2* nbr_spill;  // spill chunks (LGPU_SPILL codes each): entries 32 .. 63 of the lists longer than the table width
    int* nbr_ovf;      // by sorted slot: spill chunk of the particle (meaningful where the list is longer than M)
    int spill_cap;     // spill chunks available
    float *lambda, *density, *lambda_head;
    unsigned long long* counters;  // [0] key violations, [1] table overflows, [2] cells too crowded for the stable re-rank
};

struct lgpu_ctx {
    lgpu_config cfg;
    Geom g;
    int device;
    cudaStream_t stream;
    bool own_stream;
    int n, n_owned, n_solid, n_solid_uploaded, cap, cap_solid, M;
    int n_in, n_ghost;
    int stage_slots;      // bricks whose neighbourhood needs more stage slots re-walk the stencil (<= LGPU_STAGE_SLOTS)
    int nbY, nbX, nbZ, NB;
    int num_sms;
    int pass;             // solver pass counter of the substep being enqueued (selects the work cursor)
    int* brick_ctl;
    BrickRec* brick_rec;
    int rec_cap;
    bool generic_kernels; // test hook: run the fast-arithmetic fluid step with the generic kernels (lgpu_set_generic_kernels)
    bool grid_valid;      // cell_start/key describe the current storage
    bool solids_sorted;
    // device buffers (see View)
    float4 *pos[2], *vel[2], *pstar_unsorted, *x0, *pa, *pb;
    int *flags[2], *orig[2], *perm;
    int cur;  // which of the [2] buffers holds the current storage
    int *key_in, *rank_in, *tmp_id, *key;
    int *cell_count, *cell_start;
    unsigned long long* scan_state;
    float4 *solid_pos, *solid_pos_unsorted;
    int *solid_orig, *solid_cell_start;
    uint2* nbr16;
    int* nbr_cnt;
    uint2* nbr_spill;
    int* nbr_ovf;
    int spill_cap;
    float *lambda, *density, *lambda_head;
    unsigned long long* counters;
    float4* pstar_final;  // where the last step left x* (for dumps)
    // ---- slabs (lgpu_slab.cu) ----
    struct SlabState* slab;
    // staging
    float *h_stage, *d_stage;
    size_t stage_bytes;
    // timing
    cudaEvent_t ev[2];          // whole-step bracket (always recorded)
    cudaEvent_t ev_pool[LGPU_MAX_MARKS];  // per-launch marks (only with phase timing on)
    int ev_phase[LGPU_MAX_MARKS];
    int n_marks;
    bool phase_timing;
    long launches;
    // scratch: one growable device buffer for the per-call temporaries of the query / upload paths
    unsigned char* scratch;
    size_t scratch_bytes;
    // graphs: the launches of one substep captured once (cudaStreamBeginCapture) and replayed
    bool use_graph;
    cudaGraphExec_t graph_exec;
    long graph_sig[6];       // mode, n, n_in, n_solid, storage pointer, launches per step
    long graph_captures, graph_replays;
    lgpu_step_params last_params;
    int last_mode;  // 0 none, 1 fluid, 2 sand
};

void lgpu_set_error(const char* fmt, ...);
// at least `bytes` of device scratch (contents undefined); grows by reallocation after a stream sync
int lgpu_scratch_reserve(lgpu_ctx* c, size_t bytes);

#define CUDA_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            lgpu_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return LGPU_ERR_CUDA;                                                           \
        }                                                                                   \
    } while (0)

static inline int lgpu_blocks(long n, int block = LGPU_BLOCK) { return (int)((n + block - 1) / block); }

View lgpu_make_view(lgpu_ctx* c);

// Phase ids of lgpu_last_step_ms: 1 predict+key+histogram, 2 scan, 3 scatter+reorder, 4 neighbour
// table, 6 density/lambda kernels, 7 delta-p (fluid) or contact (sand) kernels; 5 = 6 + 7; 0 = whole step.
static inline void lgpu_mark(lgpu_ctx* c, int phase) {
    if (!c->phase_timing || c->n_marks >= LGPU_MAX_MARKS) return;
    cudaEventRecord(c->ev_pool[c->n_marks], c->stream);
    c->ev_phase[c->n_marks] = phase;
    c->n_marks++;
}

// ---- slabs ----
// Ghost refresh of slab mode (lgpu_slab.cu): after a solver kernel has written `buf`, ONE small kernel
// copies the boundary particles' new values into the neighbours' ghost slots (peer stores over
// NVLink), raises the neighbours' sequence flags and waits for theirs.  w_only: just lambda (w lane).
bool lgpu_slab_active(const lgpu_ctx* c);
bool lgpu_slab_pdl_ok(const lgpu_ctx* c);
int lgpu_slab_refresh(lgpu_ctx* c, const float4* buf, bool w_only);
int lgpu_slab_init(lgpu_ctx* c);
int lgpu_slab_check(lgpu_ctx* c);
int lgpu_preload_grid(); int lgpu_preload_neighbors(); int lgpu_preload_fluid(); int lgpu_preload_sand();
void lgpu_slab_free(lgpu_ctx* c);
struct View;
void lgpu_slab_fill_view(lgpu_ctx* c, View* v);
int lgpu_slab_begin(lgpu_ctx* c, const lgpu_step_params& p, int mode);  // predict + post halo messages
int lgpu_slab_end(lgpu_ctx* c, const lgpu_step_params& p, int mode);    // receive, grid, table, solver
int lgpu_put_sand(lgpu_ctx* c, int offset, int n, const float* pos, const float* vel, const int* flags, const int* ids);

// ---- launch wrappers, one per translation unit ----
int lgpu_launch_predict_fluid(lgpu_ctx* c, const lgpu_step_params& p);
int lgpu_launch_predict_sand(lgpu_ctx* c, const lgpu_step_params& p);
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

// programmatic dependent launch between the kernels of a captured substep (programmatic edges of the CUDA graph)
static inline bool lgpu_pdl_in_graph() {
    static const bool off = getenv("LGPU_PDL_GRAPH") && atoi(getenv("LGPU_PDL_GRAPH")) == 0;  // (LGPU_PDL_GRAPH=0: plain edges)
    return !off;
}
// whether the launches of this substep are chained programmatically (not with per-launch event marks in between;
// LGPU_PDL=0 turns it off everywhere)
static inline bool lgpu_pdl_enabled(const lgpu_ctx* c) {
    static const bool pdl_env = !(getenv("LGPU_PDL") && atoi(getenv("LGPU_PDL")) == 0);
    return pdl_env && !c->phase_timing && (!c->use_graph || lgpu_pdl_in_graph());
}
int lgpu_launch_scan_cells(lgpu_ctx* c, int* counts, int* starts, int num_cells, bool zero_counts, bool pdl);
int lgpu_launch_reorder(lgpu_ctx* c, bool reset_orig);
int lgpu_sort_solids(lgpu_ctx* c);
int lgpu_begin_passes(lgpu_ctx* c);
int lgpu_launch_build_table(lgpu_ctx* c, bool sand_order, const lgpu_step_params& p, int lambda_mode);
int lgpu_launch_dump_nbr(lgpu_ctx* c, bool sand, const long* d_off, int* d_flat);
int lgpu_launch_fluid_solver(lgpu_ctx* c, const lgpu_step_params& p);
int lgpu_launch_sand_solver(lgpu_ctx* c, const lgpu_step_params& p);

// ---------------------------------------------------------------------------------------
// Arithmetic policies.
//   Exact: every operation is a separately rounded IEEE fp32 operation in the reference's
//          order (the __f*_rn intrinsics are never contracted into FMAs), matching the
//          reference compiled with -O2 -ffp-contract=off.
//   Fast:  plain operators (nvcc contracts to FMA) and approximate reciprocal / rsqrt.
// Keys, the neighbour predicate and the contact predicate ALWAYS use Exact.
// ---------------------------------------------------------------------------------------
struct Exact {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
    static constexpr bool exact = true;
};
struct Fast {
    static __device__ __forceinline__ float add(float a, float b) { return a + b; }
    static __device__ __forceinline__ float sub(float a, float b) { return a - b; }
    static __device__ __forceinline__ float mul(float a, float b) { return a * b; }
    static __device__ __forceinline__ float div(float a, float b) { return __fdividef(a, b); }
    static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
    static constexpr bool exact = false;
};

struct F3 { float x, y, z; };
__device__ __forceinline__ F3 f3(float x, float y, float z) { F3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ F3 f3(float4 a) { return f3(a.x, a.y, a.z); }
__device__ __forceinline__ float4 f4(F3 a, float w = 0.0f) { return make_float4(a.x, a.y, a.z, w); }

// glm 0.9.9.8 semantics: dot = (x*x + y*y) + z*z on separately rounded products, length = sqrt(dot),
// normalize = v * (1/sqrt(dot)) (thirdparty/glm-0.9.9.8/glm/detail/func_geometric.inl:8-14,48-55,82-90).
template <class P> __device__ __forceinline__ F3 vadd(F3 a, F3 b) { return f3(P::add(a.x, b.x), P::add(a.y, b.y), P::add(a.z, b.z)); }
template <class P> __device__ __forceinline__ F3 vsub(F3 a, F3 b) { return f3(P::sub(a.x, b.x), P::sub(a.y, b.y), P::sub(a.z, b.z)); }
template <class P> __device__ __forceinline__ F3 vscale(F3 a, float s) { return f3(P::mul(a.x, s), P::mul(a.y, s), P::mul(a.z, s)); }
template <class P> __device__ __forceinline__ F3 vdiv(F3 a, float s) { return f3(P::div(a.x, s), P::div(a.y, s), P::div(a.z, s)); }
template <class P> __device__ __forceinline__ float vdot(F3 a, F3 b) {
    return P::add(P::add(P::mul(a.x, b.x), P::mul(a.y, b.y)), P::mul(a.z, b.z));
}
template <class P> __device__ __forceinline__ float vlen(F3 a) { return P::sqrt(vdot<P>(a, a)); }
template <class P> __device__ __forceinline__ F3 vnormalize(F3 a) { return vscale<P>(a, P::div(1.0f, P::sqrt(vdot<P>(a, a)))); }
__device__ __forceinline__ F3 vneg(F3 a) { return f3(-a.x, -a.y, -a.z); }

// Warp-aggregated `atomicAdd(ctr, 1)` for the lanes with `pred`: ONE atomic per warp instead of one per lane (hundreds of
// thousands of lanes incrementing the same counter serialise in L2: 0.1 ms per 200 k at one atomic per clock).  Called by
// all currently active lanes of the warp; returns the lane's slot, or -1 without `pred`.
__device__ __forceinline__ int warp_agg_inc(int* ctr, bool pred) {
    const unsigned act = __activemask();
    const unsigned m = __ballot_sync(act, pred);
    if (m == 0) return -1;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(ctr, __popc(m));
    base = __shfl_sync(act, base, leader);
    return pred ? base + __popc(m & ((1u << lane) - 1u)) : -1;
}

// get_cell_id, src/neighbors/Utils.hpp:24-33: IEEE division, truncation toward zero, no clamp.
// Ids outside [0, C) (undefined behaviour in the reference) are clamped and counted.
__device__ __forceinline__ int cell_x_global(const Geom& g, float x) { return __float2int_rz(__fdiv_rn(x, g.cell_size)); }
__device__ __forceinline__ int cell_id_raw(const Geom& g, F3 p) {
    int cx = __float2int_rz(__fdiv_rn(p.x, g.cell_size));
    int cy = __float2int_rz(__fdiv_rn(p.y, g.cell_size));
    int cz = __float2int_rz(__fdiv_rn(p.z, g.cell_size));
    return cy * g.gXZ + cx * g.gZ + cz;
}
// Slab mode: the key is taken in the local grid (columns x_off .. x_off + gX - 1 of the global grid,
// same y-major / z-fastest order, so the order of the particles of a slab is the global order).
// Out-of-grid coordinates are clamped per axis and counted.  `outside` = the particle is not in
// any column this context stores (only possible for solids, which are replicated and filtered).
__device__ __forceinline__ int cell_id_slab(const Geom& g, F3 p, unsigned long long* counters, bool* outside) {
    int cx = __float2int_rz(__fdiv_rn(p.x, g.cell_size)) - g.x_off;
    int cy = __float2int_rz(__fdiv_rn(p.y, g.cell_size));
    int cz = __float2int_rz(__fdiv_rn(p.z, g.cell_size));
    *outside = cx < 0 || cx >= g.gX;
    if (cy < 0 || cy >= g.gY || cz < 0 || cz >= g.gZ || *outside) {
        if (!*outside) atomicAdd(&counters[0], 1ULL);
        cx = min(max(cx, 0), g.gX - 1); cy = min(max(cy, 0), g.gY - 1); cz = min(max(cz, 0), g.gZ - 1);
    }
    return cy * g.gXZ + cx * g.gZ + cz;
}
__device__ __forceinline__ int cell_id_checked(const Geom& g, F3 p, unsigned long long* counters) {
    if (g.slab) { bool outside; return cell_id_slab(g, p, counters, &outside); }
    int id = cell_id_raw(g, p);
    if (id < 0 || id >= g.C) {
        atomicAdd(&counters[0], 1ULL);
        id = id < 0 ? 0 : g.C - 1;
    }
    return id;
}

// neighbour predicate, src/neighbors/Neighbors.cpp:347-349,433-435: dot(t,t) <= h*h, t = self - other
__device__ __forceinline__ bool within_h(const Geom& g, F3 self, F3 other) {
    F3 t = vsub<Exact>(self, other);
    return vdot<Exact>(t, t) <= g.h2;
}

// cubic_kernel, src/Kernels.cpp:6-19
template <class P> __device__ __forceinline__ float cubic_W(const Geom& g, float r) {
    float q = P::div(P::mul(r, g.kernel_factor), g.h);
    float result = 0.0f;
    if (q <= 1.0f) {
        if (q <= 0.5f) {
            float q2 = P::mul(q, q);
            float q3 = P::mul(q2, q);
            result = P::mul(g.cubic_k, P::add(P::sub(P::mul(6.0f, q3), P::mul(6.0f, q2)), 1.0f));
        } else {
            float f = P::sub(1.0f, q);
            result = P::mul(g.cubic_k, P::mul(2.0f, (float)pow((double)f, 3.0)));
        }
    }
    return result;
}

// cubic_kernel_grad, src/Kernels.cpp:26-41
template <class P> __device__ __forceinline__ F3 cubic_gradW(const Geom& g, F3 r) {
    F3 result = f3(0.0f, 0.0f, 0.0f);
    float rl = P::mul(vlen<P>(r), g.kernel_factor);
    float q = P::div(rl, g.h);
    if (rl > 1.0e-5f && q <= 1.0f) {
        F3 grad_q = vscale<P>(r, P::div(1.0f, P::mul(rl, g.h)));
        if (q <= 0.5f) {
            result = vscale<P>(grad_q, P::mul(P::mul(g.cubic_l, q), P::sub(P::mul(3.0f, q), 2.0f)));
        } else {
            float f = P::sub(1.0f, q);
            result = vscale<P>(grad_q, P::mul(g.cubic_l, P::mul(-f, f)));
        }
    }
    return result;
}

// poly6_kernel(float), src/Kernels.cpp:43-51 — std::pow(float,int) promotes: evaluated in double
__device__ __forceinline__ float poly6_W(const Geom& g, float r) {
    float result = 0.0f;
    float hf = __fmul_rn(g.h, g.kernel_factor);
    if (r <= g.h) {
        double a = 315.0f / ((double)__fmul_rn(64.0f, 3.14f) * pow((double)hf, 9.0));
        double kr = (double)__fmul_rn(g.kernel_factor, r);
        double d = (double)hf * (double)hf - kr * kr;
        result = (float)(a * (d * d * d));
    }
    return result;
}

// spiky_kernel, src/Kernels.cpp:57-67
__device__ __forceinline__ F3 spiky_gradW(const Geom& g, F3 r) {
    F3 result = f3(0.0f, 0.0f, 0.0f);
    float rl = vlen<Exact>(r);
    if (rl > 0.0f && rl <= g.h) {
        float hf = __fmul_rn(g.h, g.kernel_factor);
        double e = (double)__fsub_rn(hf, __fmul_rn(rl, g.kernel_factor));
        float temp = (float)((15.0f / ((double)3.14f * pow((double)hf, 6.0))) * (e * e));
        result = vscale<Exact>(vdiv<Exact>(r, __fmul_rn(rl, g.kernel_factor)), temp);
    }
    return result;
}

// std::pow(float,float) of s_coor (src/Simulate.cpp:8).  glibc's powf evaluates in double and
// rounds once; the device equivalent is a double pow rounded to float (n == 4: two exact-ish
// double multiplies).
__device__ __forceinline__ float powf_like_libm(float x, float n) {
    if (n == 4.0f) { double x2 = (double)x * (double)x; return (float)(x2 * x2); }
    return (float)pow((double)x, (double)n);
}
