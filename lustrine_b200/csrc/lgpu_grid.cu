// lgpu_grid.cu — predict + cell key + histogram, prefix-sum cell offsets, stable scatter/reorder.
//
// Replaces (reference paths): the integration loops src/Simulate.cpp:48-51 (fluid) and
// :188-213 / :359-390 (sand, credits), get_cell_id src/neighbors/Utils.hpp:24-33,
// Sorting::counting_sort src/neighbors/Sorting.cpp:11-33 and the gather
// src/neighbors/Neighbors.cpp:288-304.
//
// Stability.  The reference's counting sort is stable: within a cell, particles keep the order
// of their previous storage slots.  Here every particle takes an arbitrary rank inside its cell
// from the histogram atomic (k_predict_*), is scattered to cell_start[key] + rank, and
// k_reorder then recomputes its rank as the number of cell mates with a smaller reference
// slot — a deterministic (cell, reference slot) order that is bit-identical to the reference's
// permutation, without serialising anything.
#include "lgpu_internal.cuh"

// ------------------------------------------------------------------------------------------
// predict (always Exact: x* feeds the cell key, which must be bit-exact)
// ------------------------------------------------------------------------------------------
// Slab mode, end of the predict kernels: a particle whose predicted cell column left [x_lo, x_hi)
// emigrates to the neighbouring slab (record to the outbox); it now sits in this slab's ghost column,
// so its slot lives on as a ghost copy that the new owner refreshes like any other ghost.  A particle
// in the first / last gw owned columns is sent as a ghost copy.  Returns true if the particle emigrated.
__device__ __forceinline__ bool slab_classify(const View& v, int i, F3 x, F3 vel, F3 xs, int flags) {
    const int cxg = cell_x_global(v.g, xs.x);
    const int side = cxg < v.g.x_lo ? 0 : (cxg >= v.g.x_hi ? 1 : -1);
    HaloRec rec;
    rec.pos = f4(x); rec.vel = f4(vel); rec.pstar = f4(xs); rec.flags = flags; rec.orig = v.orig_in[i]; rec.pad0 = rec.pad1 = 0;
    // (outbox slots are handed out per warp, see warp_agg_inc; the order of the records is the message order both ends use)
    bool emigrated = false;
#pragma unroll
    for (int s = 0; s < 2; s++) {
        const bool mig = side == s && v.has_nbr[s];
        const int slot = warp_agg_inc(&v.out_cnt[s], mig);
        if (mig) {
            if (slot < v.halo_cap) { v.out_mig[s][slot] = rec; v.mig_src[s][slot] = i; }
            else atomicAdd(&v.out_cnt[4], 1);
            emigrated = true;
        }
    }
#pragma unroll
    for (int s = 0; s < 2; s++) {
        const bool edge = s == 0 ? cxg < v.g.x_lo + v.g.gw : cxg >= v.g.x_hi - v.g.gw;  // the first / last gw owned columns
        const bool gho = !emigrated && edge && v.has_nbr[s];
        const int slot = warp_agg_inc(&v.out_cnt[2 + s], gho);
        if (gho) {
            if (slot < v.halo_cap) { v.out_gho[s][slot] = rec; v.gho_src[s][slot] = i; }
            else atomicAdd(&v.out_cnt[4], 1);
        }
    }
    // An open side (no neighbour, and the plan stops short of the domain wall: the boundaries were cropped to the occupied
    // columns so that the empty part of the domain costs no cells): count the particles that approach the end of the
    // planned columns (the host re-plans) and those that have left them (the step fails).
#pragma unroll
    for (int s = 0; s < 2; s++) {
        if (v.has_nbr[s] || (s == 0 ? v.g.x_lo <= 0 : v.g.x_hi >= v.g.gXg)) continue;  // (uniform)
        const bool left = s == 0 ? cxg < v.g.x_lo : cxg >= v.g.x_hi;
        const bool near = left || (s == 0 ? cxg < v.g.x_lo + v.g.guard : cxg >= v.g.x_hi - v.g.guard);
        warp_agg_inc(&v.out_cnt[5], near);
        warp_agg_inc(&v.out_cnt[6], left);
    }
    return emigrated;
}

__global__ void __launch_bounds__(LGPU_BLOCK) k_predict_fluid(View v, float dt, F3 gm /* gravity*mass */) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.n_in) return;
    int key;
    const int a = v.g.slab ? v.flags_in[i] : 0;
    if (a & (LGPU_FLAG_DEAD | LGPU_FLAG_GHOST)) {
        // slab mode: an emigrated particle or last step's ghost copy — dropped by the sort (trash cell C)
        v.flags_in[i] = LGPU_FLAG_DEAD;
        key = v.g.C;
    } else {
        F3 x = f3(v.pos_in[i]);
        F3 vel = f3(v.vel_in[i]);
        // src/Simulate.cpp:49-50: v += gravity * mass * dt;  x* = x + v * dt
        vel = vadd<Exact>(vel, vscale<Exact>(gm, dt));
        F3 xs = vadd<Exact>(x, vscale<Exact>(vel, dt));
        // (the predicted velocity is not stored: the last solver pass rewrites v from the positions, src/Simulate.cpp:110)
        v.pstar_in[i] = f4(xs);
        if (v.g.slab && slab_classify(v, i, x, vel, xs, a)) v.flags_in[i] = a | LGPU_FLAG_GHOST;
        key = cell_id_checked(v.g, xs, v.counters);
    }
    v.key_in[i] = key;
    // (the dead slots of a slab — last step's ghosts, the emigrants — all land in the trash cell: one atomic per warp)
    const int r_dead = v.g.slab ? warp_agg_inc(&v.cell_count[v.g.C], key == v.g.C) : -1;
    v.rank_in[i] = key == v.g.C ? r_dead : atomicAdd(&v.cell_count[key], 1);
}

struct SandPredict {
    float dt;
    F3 gravity, player;
    int attract_flag, blow_flag, prev_attract_flag, credits;
    float attract_radius, blow_radius, attract_coeff, blow_coeff, mass;
};

__global__ void __launch_bounds__(LGPU_BLOCK) k_predict_sand(View v, SandPredict s) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.n_in) return;
    typedef Exact P;
    int key;
    int a = v.flags_in[i];
    if (v.g.slab && (a & (LGPU_FLAG_DEAD | LGPU_FLAG_GHOST))) {
        v.flags_in[i] = LGPU_FLAG_DEAD;
        key = v.g.C;
    } else {
        F3 x = f3(v.pos_in[i]);
        F3 vel = f3(v.vel_in[i]);
        F3 xs;
        const float r = v.g.radius;
        float w = P::div(1.0f, s.mass);                                         // :189
        if (s.credits && (a & 2)) vel = f3(0.0f, -1.0f, 0.0f);                   // :362-364
        vel = vadd<P>(vel, vscale<P>(s.gravity, s.dt));                          // :190
        if (s.prev_attract_flag && !s.attract_flag) a = s.credits ? (a & ~1) : 0;  // :191-193
        if (s.attract_flag) {
            if (vlen<P>(vsub<P>(s.player, x)) < s.attract_radius) a = s.credits ? ((a | 1) & ~2) : 1;  // :195-197
            if (s.credits ? (a & 1) : a) {                                       // :198-201
                F3 to = vsub<P>(vadd<P>(s.player, f3(0.0f, 1.5f, 0.0f)), x);
                F3 t = vscale<P>(vnormalize<P>(to), 1.0f);
                t = vscale<P>(t, s.attract_coeff); t = vscale<P>(t, r); t = vscale<P>(t, s.dt); t = vscale<P>(t, w);
                vel = vadd<P>(vel, t);
            }
        }
        if (s.blow_flag) {                                                        // :203-208
            F3 to = vsub<P>(s.player, x);
            float len = vlen<P>(to);
            if (len < s.blow_radius) {
                float bk = len > s.blow_radius ? 0.0f : P::sub(1.0f, P::div(len, s.blow_radius));  // :143-154
                F3 t = vscale<P>(vneg(vnormalize<P>(to)), bk);
                t = vscale<P>(t, s.blow_coeff); t = vscale<P>(t, r); t = vscale<P>(t, w);
                vel = vadd<P>(vel, t);
                if (s.credits) a &= ~2;
            }
        }
        xs = vadd<P>(x, vscale<P>(vel, s.dt));                                   // :210
        // :212 glm::clamp(x*, r, D - r) = min(max(x, lo), hi)
        xs.x = fminf(fmaxf(xs.x, r), P::sub(v.g.domainX, r));
        xs.y = fminf(fmaxf(xs.y, r), P::sub(v.g.domainY, r));
        xs.z = fminf(fmaxf(xs.z, r), P::sub(v.g.domainZ, r));
        v.flags_in[i] = a;  // (the predicted velocity is not stored: the last contact pass rewrites v, src/Simulate.cpp:317)
        v.pstar_in[i] = f4(xs);
        if (v.g.slab && slab_classify(v, i, x, vel, xs, a)) v.flags_in[i] = a | LGPU_FLAG_GHOST;
        key = cell_id_checked(v.g, xs, v.counters);
    }
    v.key_in[i] = key;
    // (the dead slots of a slab — last step's ghosts, the emigrants — all land in the trash cell: one atomic per warp)
    const int r_dead = v.g.slab ? warp_agg_inc(&v.cell_count[v.g.C], key == v.g.C) : -1;
    v.rank_in[i] = key == v.g.C ? r_dead : atomicAdd(&v.cell_count[key], 1);
}

int lgpu_launch_predict_fluid(lgpu_ctx* c, const lgpu_step_params& p) {
    if (c->n_in == 0) return LGPU_OK;
    View v = lgpu_make_view(c);
    float dt = fminf(fmaxf(p.dt, 0.001f), 0.01f);  // src/Simulate.cpp:31
    // gravity * mass evaluated once on the host in fp32, like (gravity * mass) in :49
    F3 gm;
    gm.x = p.gravity[0] * p.mass; gm.y = p.gravity[1] * p.mass; gm.z = p.gravity[2] * p.mass;
    k_predict_fluid<<<lgpu_blocks(c->n_in), LGPU_BLOCK, 0, c->stream>>>(v, dt, gm);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return LGPU_OK;
}

int lgpu_launch_predict_sand(lgpu_ctx* c, const lgpu_step_params& p) {
    if (c->n_in == 0) return LGPU_OK;
    View v = lgpu_make_view(c);
    SandPredict s;
    s.dt = p.dt;
    s.gravity.x = p.gravity[0]; s.gravity.y = p.gravity[1]; s.gravity.z = p.gravity[2];
    s.player.x = p.player_position[0]; s.player.y = p.player_position[1]; s.player.z = p.player_position[2];
    s.attract_flag = p.attract_flag; s.blow_flag = p.blow_flag; s.prev_attract_flag = p.prev_attract_flag;
    s.credits = p.credits;
    s.attract_radius = p.attract_radius; s.blow_radius = p.blow_radius;
    s.attract_coeff = p.attract_coeff; s.blow_coeff = p.blow_coeff; s.mass = p.mass;
    k_predict_sand<<<lgpu_blocks(c->n_in), LGPU_BLOCK, 0, c->stream>>>(v, s);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return LGPU_OK;
}

// ------------------------------------------------------------------------------------------
// exclusive prefix sum over the cell histogram (inclusive scan of Sorting::counting_sort,
// src/neighbors/Sorting.cpp:22-24, as cell offsets): ONE pass over the grid.  Reads the histogram once
// (int4, coalesced), writes the offsets once and zeroes the histogram for the next substep in the same
// pass: 12 bytes per cell.
// Tiles of 8192 cells, one 512-thread block each, taken in ticket order.  A tile publishes its aggregate
// and then looks back over its predecessors a WINDOW of 512 tiles at a time, one status word per thread
// (the usual warp-wide look-back walks 32 tiles per round trip to L2: 25 dependent rounds for the 400 tiles
// of the 1 M dam break; here the whole history is one round): the nearest predecessor that already knows
// its inclusive prefix ends the walk.
// ------------------------------------------------------------------------------------------
#define SCAN_THREADS 512
#define SCAN_WARPS (SCAN_THREADS / 32)
#define SCAN_SUB 4                                // sub-tiles of 512 threads x int4 per tile
#define SCAN_TILE (SCAN_SUB * SCAN_THREADS * 4)   // 8192 cells: the 3.3 M cells of the 1 M dam break are 401 blocks, three per SM = one wave

__device__ __forceinline__ int warp_incl_scan(int x) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) >= o) x += y;
    }
    return x;
}

// Scan state (device memory, zeroed once when it is allocated): [0] ticket counter, [1] tiles finished, [2] epoch,
// [3 + tile] status word of the tile: bits 62..63 = 1 tile aggregate / 2 inclusive prefix, bits 32..61 = epoch of the
// launch that wrote it, low 32 bits = value.  A word of another epoch reads as "not ready", so nothing has to be
// cleared between launches: the last tile to finish resets the counters and bumps the epoch (launches on the stream
// are serialised, and the same captured graph node can be replayed).
#define SCAN_AGG 1ULL
#define SCAN_PFX 2ULL
#define SCAN_EPOCH_MASK 0x3fffffffULL
__device__ __forceinline__ unsigned long long scan_word(unsigned long long flag, unsigned long long epoch, int value) {
    return (flag << 62) | (epoch << 32) | (unsigned int)value;
}

__global__ void __launch_bounds__(SCAN_THREADS, 3) k_scan_cells(int* __restrict__ counts, int n, int* __restrict__ starts,
                                                             unsigned long long* __restrict__ state, int zero_counts) {
    __shared__ int warp_sums[SCAN_SUB][SCAN_WARPS];
    __shared__ int look_sum[SCAN_WARPS], look_end[SCAN_WARPS];
    __shared__ int s_tile, s_prefix;
    __shared__ unsigned long long s_epoch;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int nb = gridDim.x;
    volatile unsigned long long* vs = state + 3;
    pdl_trigger();
    pdl_wait();  // (a no-op unless the launch is chained programmatically to its predecessor)
    if (tid == 0) {
        s_epoch = *(volatile unsigned long long*)(state + 2) & SCAN_EPOCH_MASK;
        s_tile = (int)atomicAdd(&state[0], 1ULL);  // ticket: a tile only ever waits for tiles that started before it
    }
    __syncthreads();
    const int tile = s_tile;
    const unsigned long long epoch = s_epoch;
    const long base = (long)tile * SCAN_TILE;
    int4 vals[SCAN_SUB];
    int tsum[SCAN_SUB], inc[SCAN_SUB];
#pragma unroll
    for (int k = 0; k < SCAN_SUB; k++) {
        const long idx = base + (long)k * (SCAN_THREADS * 4) + tid * 4;
        int4 a = make_int4(0, 0, 0, 0);
        if (idx + 3 < n) a = *reinterpret_cast<const int4*>(counts + idx);
        else {
            if (idx < n) a.x = counts[idx];
            if (idx + 1 < n) a.y = counts[idx + 1];
            if (idx + 2 < n) a.z = counts[idx + 2];
        }
        vals[k] = a;
        tsum[k] = a.x + a.y + a.z + a.w;
    }
#pragma unroll
    for (int k = 0; k < SCAN_SUB; k++) inc[k] = warp_incl_scan(tsum[k]);
    if (lane == 31) {
#pragma unroll
        for (int k = 0; k < SCAN_SUB; k++) warp_sums[k][w] = inc[k];
    }
    __syncthreads();
    // every warp scans the SCAN_SUB x 32 warp sums for itself: offsets of its own warp and the tile aggregate
    int woff[SCAN_SUB];
    int agg = 0;
#pragma unroll
    for (int k = 0; k < SCAN_SUB; k++) {
        const int ws = lane < SCAN_WARPS ? warp_sums[k][lane] : 0;
        const int wi = warp_incl_scan(ws);
        woff[k] = agg + __shfl_sync(0xffffffffu, wi - ws, w);
        agg += __shfl_sync(0xffffffffu, wi, 31);
    }
    if (tile == 0) {
        if (tid == 0) { vs[0] = scan_word(SCAN_PFX, epoch, agg); s_prefix = 0; }
        __syncthreads();
    } else {
        if (tid == 0) vs[tile] = scan_word(SCAN_AGG, epoch, agg);
        int running = 0;
        for (int hi = tile - 1;; hi -= SCAN_THREADS) {  // window: tiles hi, hi - 1, ..., hi - 511 (thread t looks at hi - t)
            const int look = hi - tid;
            unsigned long long st = scan_word(SCAN_PFX, epoch, 0);  // (before tile 0: prefix 0)
            if (look >= 0) {
                const long long t0 = clock64();
                do {
                    st = vs[look];
                    // (a predecessor that never publishes would be a bug or a killed launch: give up after ~2 s instead of hanging the device)
                    if (clock64() - t0 > 4000000000LL) { st = scan_word(SCAN_PFX, epoch, 0); break; }
                } while ((st >> 62) == 0 || ((st >> 32) & SCAN_EPOCH_MASK) != epoch);
            }
            const unsigned pm = __ballot_sync(0xffffffffu, (st >> 62) == SCAN_PFX);
            const int first = pm ? __ffs(pm) - 1 : 32;
            int contrib = lane <= first ? (int)(unsigned int)st : 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
            if (lane == 0) { look_sum[w] = contrib; look_end[w] = pm ? 1 : 0; }
            __syncthreads();
            // warps in window order until the first one that met an inclusive prefix
            const unsigned em = __ballot_sync(0xffffffffu, lane < SCAN_WARPS && look_end[lane] != 0);
            const int wfirst = em ? __ffs(em) - 1 : 32;
            int part = (lane <= wfirst && lane < SCAN_WARPS) ? look_sum[lane] : 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            running += part;
            __syncthreads();  // (look_sum / look_end are rewritten by the next window)
            if (em) break;
        }
        if (tid == 0) { vs[tile] = scan_word(SCAN_PFX, epoch, running + agg); s_prefix = running; }
        __syncthreads();
    }
    const int carry = s_prefix;
#pragma unroll
    for (int k = 0; k < SCAN_SUB; k++) {
        const int ex = carry + woff[k] + inc[k] - tsum[k];
        const long idx = base + (long)k * (SCAN_THREADS * 4) + tid * 4;
        const int4 a = vals[k];
        const int4 o4 = make_int4(ex, ex + a.x, ex + a.x + a.y, ex + a.x + a.y + a.z);
        if (idx + 3 < n) {
            *reinterpret_cast<int4*>(starts + idx) = o4;
            if (zero_counts) *reinterpret_cast<int4*>(counts + idx) = make_int4(0, 0, 0, 0);
        } else {
            if (idx < n) { starts[idx] = o4.x; if (zero_counts) counts[idx] = 0; }
            if (idx + 1 < n) { starts[idx + 1] = o4.y; if (zero_counts) counts[idx + 1] = 0; }
            if (idx + 2 < n) { starts[idx + 2] = o4.z; if (zero_counts) counts[idx + 2] = 0; }
        }
        // starts[n] = grand total: written by the thread that owns element n-1
        if (n - 1 >= idx && n - 1 <= idx + 3) starts[n] = ex + tsum[k];
    }
    // the last tile to get here (every tile has read all the status words it needs by now) re-arms the state
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(&state[1], 1ULL) == (unsigned long long)(nb - 1)) {
            state[0] = 0; state[1] = 0; state[2] = (epoch + 1) & SCAN_EPOCH_MASK;
        }
    }
}

int lgpu_launch_scan_cells(lgpu_ctx* c, int* counts, int* starts, int num_cells, bool zero_counts, bool pdl) {
    int nb = (num_cells + SCAN_TILE - 1) / SCAN_TILE;
    CUDA_TRY(launch_pdl(k_scan_cells, nb, SCAN_THREADS, 0, c->stream, pdl, counts, num_cells, starts, c->scan_state, zero_counts ? 1 : 0));
    c->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return LGPU_OK;
}

// ------------------------------------------------------------------------------------------
// scatter + stable reorder
// ------------------------------------------------------------------------------------------
// The scattered record of a particle: {reference slot, unsorted index, cell key}.  k_reorder ranks a particle among
// its cell mates from the records next to its own (one contiguous read) instead of chasing index -> reference slot
// per mate, and needs no further look-up to find its cell.
__global__ void __launch_bounds__(LGPU_BLOCK) k_scatter_ids(View v) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_trigger();
    pdl_wait();
    if (i >= v.n_in) return;
    const int key = v.key_in[i];
    v.sort_rec[v.cell_start[key] + v.rank_in[i]] = make_int4(v.orig_in[i], i, key, 0);
}

__global__ void __launch_bounds__(LGPU_BLOCK) k_reorder(View v, int reset_orig) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_trigger();
    pdl_wait();
    if (s >= v.n_in) return;
    const int4 rec = v.sort_rec[s];
    const int mine = rec.x, i = rec.y, c = rec.z;
    if (c == v.g.C) return;  // slab mode: dead slot (trash cell), dropped
    // (everything the particle brings along is requested before the rank loop: one round of memory latency)
    const float4 x = v.pos_in[i];
    float4 ps = v.pstar_in[i];
    const int fl = v.flags_in[i];
    const int b = v.cell_start[c], e = v.cell_start[c + 1];
    int r;
    if (e - b <= LGPU_STABLE_MAX) {
        // stable order inside the cell (ascending reference slot, src/neighbors/Sorting.cpp:26-33): rank among the cell's members
        r = 0;
        for (int u = b; u < e; u++) r += (v.sort_rec[u].x < mine) ? 1 : 0;
    } else {
        // More members than any packing puts into one cell: particles whose coordinates left the grid and were clamped
        // into a border cell (undefined behaviour in the reference).  The quadratic re-rank is skipped;
        // the cell keeps the order in which its members arrived (a valid permutation, not the reference's), counted.
        r = s - b;
        if (s == b) atomicAdd(&v.counters[2], 1ULL);
    }
    const int dst = b + r;
    v.pos[dst] = x;
    // (the sorted copy of the velocities is not kept: both solvers recompute v from the positions at the end of the
    // step, src/Simulate.cpp:110,317)
    // the w lane of a predicted position carries the particle's own sorted slot: the sand solver gets the
    // storage slot of a staged neighbour from it (the fluid solver overwrites it with lambda)
    ps.w = __int_as_float(dst);
    v.x0[dst] = ps;
    v.flags[dst] = fl;
    v.key[dst] = c;
    v.perm[dst] = mine;
    // the reference permutes its own storage in the sand path (src/neighbors/Neighbors.cpp:296-300):
    // the sorted slot becomes the particle's reference slot.  The fluid path keeps its storage order.
    v.orig[dst] = reset_orig ? dst : mine;
    if (v.g.slab) v.inv[i] = dst;
}

int lgpu_launch_reorder(lgpu_ctx* c, bool reset_orig) {
    if (c->n_in == 0) return LGPU_OK;
    View v = lgpu_make_view(c);
    // chained programmatically to the scan (and k_reorder to k_scatter_ids) on a single GPU; slab mode has plain
    // launches in between
    const bool pdl = lgpu_pdl_enabled(c) && !c->g.slab;
    CUDA_TRY(launch_pdl(k_scatter_ids, lgpu_blocks(c->n_in), LGPU_BLOCK, 0, c->stream, pdl, v));
    CUDA_TRY(launch_pdl(k_reorder, lgpu_blocks(c->n_in), LGPU_BLOCK, 0, c->stream, pdl, v, reset_orig ? 1 : 0));
    c->launches += 2;
    CUDA_TRY(cudaGetLastError());
    return LGPU_OK;
}

// ------------------------------------------------------------------------------------------
// solids: binned once, ascending upload index inside each cell (src/neighbors/Neighbors.cpp:266-272)
// ------------------------------------------------------------------------------------------
__global__ void k_solid_keys(Geom g, const float4* pos, int n, int* keys, int* ranks, int* counts, unsigned long long* counters) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int key;
    if (g.slab) {
        // solids are replicated on every slab; each keeps those inside its columns (ghost columns included)
        bool outside;
        key = cell_id_slab(g, f3(pos[i]), counters, &outside);
        if (outside) key = g.C;
    } else {
        key = cell_id_checked(g, f3(pos[i]), counters);
    }
    keys[i] = key;
    ranks[i] = atomicAdd(&counts[key], 1);
}
__global__ void k_solid_scatter(const int* keys, const int* ranks, const int* starts, int n, int* tmp) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    tmp[starts[keys[i]] + ranks[i]] = i;
}
__global__ void k_solid_reorder(const float4* pos_in, const int* keys, const int* starts, const int* tmp, int n, float4* pos_out, int* orig_out, int trash) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int i = tmp[s];
    int c = keys[i];
    if (c == trash) return;
    int b = starts[c], e = starts[c + 1];
    int r = 0;
    for (int u = b; u < e; u++) r += (tmp[u] < i) ? 1 : 0;
    pos_out[b + r] = pos_in[i];
    orig_out[b + r] = i;
}

int lgpu_sort_solids(lgpu_ctx* c) {
    const int n = c->n_solid_uploaded, C = c->g.C;
    c->n_solid = n;
    CUDA_TRY(cudaMemsetAsync(c->solid_cell_start, 0, sizeof(int) * ((size_t)C + 2), c->stream));
    if (n > 0) {
        const size_t n4 = ((size_t)n + 3) & ~(size_t)3;
        int st0 = lgpu_scratch_reserve(c, sizeof(int) * (3 * n4 + (size_t)C + 2));
        if (st0) return st0;
        int* keys = (int*)c->scratch;
        int *ranks = keys + n4, *tmp = ranks + n4, *counts = tmp + n4;  // (counts is 16-byte aligned: the scan loads int4)
        CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(int) * ((size_t)C + 2), c->stream));
        k_solid_keys<<<lgpu_blocks(n), LGPU_BLOCK, 0, c->stream>>>(c->g, c->solid_pos_unsorted, n, keys, ranks, counts, c->counters);
        int st = lgpu_launch_scan_cells(c, counts, c->solid_cell_start, C + 1, false, false);
        if (st) return st;
        k_solid_scatter<<<lgpu_blocks(n), LGPU_BLOCK, 0, c->stream>>>(keys, ranks, c->solid_cell_start, n, tmp);
        k_solid_reorder<<<lgpu_blocks(n), LGPU_BLOCK, 0, c->stream>>>(c->solid_pos_unsorted, keys, c->solid_cell_start, tmp, n, c->solid_pos, c->solid_orig, C);
        c->launches += 3;
        CUDA_TRY(cudaGetLastError());
        int kept = n;
        CUDA_TRY(cudaMemcpyAsync(&kept, c->solid_cell_start + C, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        c->n_solid = kept;  // slab mode: the solids inside this context's columns
    }
    c->solids_sorted = true;
    return LGPU_OK;
}

// ------------------------------------------------------------------------------------------
// stand-alone stable counting sort on caller keys (known-answer test of the reference's
// experiments/unit_tests/main.cpp:46-89)
// ------------------------------------------------------------------------------------------
__global__ void k_cs_hist(const int* keys, int n, int* ranks, int* counts) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ranks[i] = atomicAdd(&counts[keys[i]], 1);
}
__global__ void k_cs_rank(const int* keys, const int* starts, const int* tmp, int n, int* sorted) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int i = tmp[s];
    int c = keys[i];
    int b = starts[c], e = starts[c + 1];
    int r = 0;
    for (int u = b; u < e; u++) r += (tmp[u] < i) ? 1 : 0;
    sorted[b + r] = i;
}

int lgpu_counting_sort(const int* keys, int n, int num_cells, int* sorted, int device) {
    if (n < 0 || num_cells <= 0 || (!keys && n) || (!sorted && n)) return LGPU_ERR_ARG;
    if (device >= 0) CUDA_TRY(cudaSetDevice(device));
    if (n == 0) return LGPU_OK;
    for (int i = 0; i < n; i++) if (keys[i] < 0 || keys[i] >= num_cells) return LGPU_ERR_ARG;
    lgpu_ctx tmpctx = {};
    CUDA_TRY(cudaStreamCreateWithFlags(&tmpctx.stream, cudaStreamNonBlocking));
    int *d_keys, *d_ranks, *d_counts, *d_starts, *d_tmp, *d_sorted;
    int nb = (num_cells + SCAN_TILE - 1) / SCAN_TILE;
    CUDA_TRY(cudaMalloc(&d_keys, sizeof(int) * n));
    CUDA_TRY(cudaMalloc(&d_ranks, sizeof(int) * n));
    CUDA_TRY(cudaMalloc(&d_tmp, sizeof(int) * n));
    CUDA_TRY(cudaMalloc(&d_sorted, sizeof(int) * n));
    CUDA_TRY(cudaMalloc(&d_counts, sizeof(int) * ((size_t)num_cells + 1)));
    CUDA_TRY(cudaMalloc(&d_starts, sizeof(int) * ((size_t)num_cells + 1)));
    CUDA_TRY(cudaMalloc(&tmpctx.scan_state, sizeof(unsigned long long) * ((size_t)nb + 4)));
    CUDA_TRY(cudaMemsetAsync(tmpctx.scan_state, 0, sizeof(unsigned long long) * ((size_t)nb + 4), tmpctx.stream));
    CUDA_TRY(cudaMemcpyAsync(d_keys, keys, sizeof(int) * n, cudaMemcpyHostToDevice, tmpctx.stream));
    CUDA_TRY(cudaMemsetAsync(d_counts, 0, sizeof(int) * ((size_t)num_cells + 1), tmpctx.stream));
    k_cs_hist<<<lgpu_blocks(n), LGPU_BLOCK, 0, tmpctx.stream>>>(d_keys, n, d_ranks, d_counts);
    int st = lgpu_launch_scan_cells(&tmpctx, d_counts, d_starts, num_cells, false, false);
    if (st) return st;
    k_solid_scatter<<<lgpu_blocks(n), LGPU_BLOCK, 0, tmpctx.stream>>>(d_keys, d_ranks, d_starts, n, d_tmp);
    k_cs_rank<<<lgpu_blocks(n), LGPU_BLOCK, 0, tmpctx.stream>>>(d_keys, d_starts, d_tmp, n, d_sorted);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(sorted, d_sorted, sizeof(int) * n, cudaMemcpyDeviceToHost, tmpctx.stream));
    CUDA_TRY(cudaStreamSynchronize(tmpctx.stream));
    cudaFree(d_keys); cudaFree(d_ranks); cudaFree(d_tmp); cudaFree(d_sorted); cudaFree(d_counts); cudaFree(d_starts);
    cudaFree(tmpctx.scan_state);
    cudaStreamDestroy(tmpctx.stream);
    return LGPU_OK;
}


// CUDA loads kernels lazily at their first launch, and that load can wait for the device to go
// idle.  In slab mode a context may sit in a flag-wait kernel until its neighbour launches a
// kernel for the first time — so every kernel of the step is loaded when the context is created.
#define LGPU_PRELOAD(f) do { cudaFuncAttributes a; CUDA_TRY(cudaFuncGetAttributes(&a, f)); } while (0)
int lgpu_preload_grid() {
    LGPU_PRELOAD(k_predict_fluid); LGPU_PRELOAD(k_predict_sand); LGPU_PRELOAD(k_scan_cells);
    LGPU_PRELOAD(k_scatter_ids); LGPU_PRELOAD(k_reorder);
    LGPU_PRELOAD(k_solid_keys); LGPU_PRELOAD(k_solid_scatter); LGPU_PRELOAD(k_solid_reorder);
    return LGPU_OK;
}
