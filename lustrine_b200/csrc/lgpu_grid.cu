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
__global__ void __launch_bounds__(LGPU_BLOCK) k_predict_fluid(View v, float dt, F3 gm /* gravity*mass */) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.n) return;
    F3 x = f3(v.pos_in[i]);
    F3 vel = f3(v.vel_in[i]);
    F3 xs;
    if (i < v.n_owned) {
        // src/Simulate.cpp:49-50: v += gravity * mass * dt;  x* = x + v * dt
        vel = vadd<Exact>(vel, vscale<Exact>(gm, dt));
        xs = vadd<Exact>(x, vscale<Exact>(vel, dt));
        v.vel_in[i] = f4(vel);
    } else {
        xs = f3(v.pstar_in[i]);  // ghost: predicted by its owner
    }
    v.pstar_in[i] = f4(xs);
    int key = cell_id_checked(v.g, xs, v.counters);
    v.key_in[i] = key;
    v.rank_in[i] = atomicAdd(&v.cell_count[key], 1);
}

struct SandPredict {
    float dt;
    F3 gravity, player;
    int attract_flag, blow_flag, prev_attract_flag, credits;
    float attract_radius, blow_radius, attract_coeff, blow_coeff, mass;
};

__global__ void __launch_bounds__(LGPU_BLOCK) k_predict_sand(View v, SandPredict s) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.n) return;
    typedef Exact P;
    F3 x = f3(v.pos_in[i]);
    F3 vel = f3(v.vel_in[i]);
    F3 xs;
    if (i < v.n_owned) {
        int a = v.flags_in[i];
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
        v.vel_in[i] = f4(vel);
        v.flags_in[i] = a;
    } else {
        xs = f3(v.pstar_in[i]);
    }
    v.pstar_in[i] = f4(xs);
    int key = cell_id_checked(v.g, xs, v.counters);
    v.key_in[i] = key;
    v.rank_in[i] = atomicAdd(&v.cell_count[key], 1);
}

int lgpu_launch_predict_fluid(lgpu_ctx* c, const lgpu_step_params& p) {
    if (c->n == 0) return LGPU_OK;
    View v = lgpu_make_view(c);
    float dt = fminf(fmaxf(p.dt, 0.001f), 0.01f);  // src/Simulate.cpp:31
    // gravity * mass evaluated once on the host in fp32, like (gravity * mass) in :49
    F3 gm;
    gm.x = p.gravity[0] * p.mass; gm.y = p.gravity[1] * p.mass; gm.z = p.gravity[2] * p.mass;
    k_predict_fluid<<<lgpu_blocks(c->n), LGPU_BLOCK, 0, c->stream>>>(v, dt, gm);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return LGPU_OK;
}

int lgpu_launch_predict_sand(lgpu_ctx* c, const lgpu_step_params& p) {
    if (c->n == 0) return LGPU_OK;
    View v = lgpu_make_view(c);
    SandPredict s;
    s.dt = p.dt;
    s.gravity.x = p.gravity[0]; s.gravity.y = p.gravity[1]; s.gravity.z = p.gravity[2];
    s.player.x = p.player_position[0]; s.player.y = p.player_position[1]; s.player.z = p.player_position[2];
    s.attract_flag = p.attract_flag; s.blow_flag = p.blow_flag; s.prev_attract_flag = p.prev_attract_flag;
    s.credits = p.credits;
    s.attract_radius = p.attract_radius; s.blow_radius = p.blow_radius;
    s.attract_coeff = p.attract_coeff; s.blow_coeff = p.blow_coeff; s.mass = p.mass;
    k_predict_sand<<<lgpu_blocks(c->n), LGPU_BLOCK, 0, c->stream>>>(v, s);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return LGPU_OK;
}

// ------------------------------------------------------------------------------------------
// exclusive prefix sum over the cell histogram: three phases, 4096 cells per block
// ------------------------------------------------------------------------------------------
#define SCAN_THREADS 256
#define SCAN_ITEMS 16
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ int warp_incl_scan(int x) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) >= o) x += y;
    }
    return x;
}

// block-wide exclusive scan of one int per thread; returns the exclusive prefix, total in *total
__device__ __forceinline__ int block_excl_scan(int x, int* total) {
    __shared__ int warp_sums[SCAN_THREADS / 32];
    __shared__ int block_total;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = warp_incl_scan(x);
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    if (w == 0) {
        int s = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
        int si = warp_incl_scan(s);
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = si - s;
        if (lane == SCAN_THREADS / 32 - 1) block_total = si;
    }
    __syncthreads();
    int r = inc - x + warp_sums[w];
    *total = block_total;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const int* __restrict__ counts, int n, int* __restrict__ block_sums) {
    long base = (long)blockIdx.x * SCAN_TILE;
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        long idx = base + (long)k * SCAN_THREADS + threadIdx.x;
        if (idx < n) s += counts[idx];
    }
    int total;
    block_excl_scan(s, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: exclusive scan of the block sums (any count, looped)
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_block_sums(int* block_sums, int nb) {
    int carry = 0;
    for (int base = 0; base < nb; base += SCAN_THREADS) {
        int idx = base + threadIdx.x;
        int x = idx < nb ? block_sums[idx] : 0;
        int total;
        int ex = block_excl_scan(x, &total);
        if (idx < nb) block_sums[idx] = ex + carry;
        carry += total;
    }
    if (threadIdx.x == 0) block_sums[nb] = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(int* __restrict__ counts, int n, const int* __restrict__ block_sums,
                                                             int* __restrict__ starts, int zero_counts) {
    // each thread owns SCAN_ITEMS consecutive cells -> one serial scan + one block scan
    long base = (long)blockIdx.x * SCAN_TILE + (long)threadIdx.x * SCAN_ITEMS;
    int vals[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        long idx = base + k;
        vals[k] = idx < n ? counts[idx] : 0;
        s += vals[k];
    }
    int total;
    int ex = block_excl_scan(s, &total) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        long idx = base + k;
        if (idx < n) {
            starts[idx] = ex;
            if (zero_counts) counts[idx] = 0;
        }
        ex += vals[k];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1) starts[n] = block_sums[gridDim.x];
}

int lgpu_launch_scan_cells(lgpu_ctx* c, int* counts, int* starts, int num_cells, bool zero_counts) {
    int nb = (num_cells + SCAN_TILE - 1) / SCAN_TILE;
    k_scan_reduce<<<nb, SCAN_THREADS, 0, c->stream>>>(counts, num_cells, c->scan_block_sums);
    k_scan_block_sums<<<1, SCAN_THREADS, 0, c->stream>>>(c->scan_block_sums, nb);
    k_scan_apply<<<nb, SCAN_THREADS, 0, c->stream>>>(counts, num_cells, c->scan_block_sums, starts, zero_counts ? 1 : 0);
    c->launches += 3;
    CUDA_TRY(cudaGetLastError());
    return LGPU_OK;
}

// ------------------------------------------------------------------------------------------
// scatter + stable reorder
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LGPU_BLOCK) k_scatter_ids(View v) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.n) return;
    v.tmp_id[v.cell_start[v.key_in[i]] + v.rank_in[i]] = i;
}

__global__ void __launch_bounds__(LGPU_BLOCK) k_reorder(View v, int reset_orig) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= v.n) return;
    int i = v.tmp_id[s];
    int c = v.key_in[i];
    int b = v.cell_start[c], e = v.cell_start[c + 1];
    int mine = v.orig_in[i];
    int r = 0;
    for (int u = b; u < e; u++) r += (v.orig_in[v.tmp_id[u]] < mine) ? 1 : 0;
    int dst = b + r;
    v.pos[dst] = v.pos_in[i];
    v.vel[dst] = v.vel_in[i];
    v.x0[dst] = v.pstar_in[i];
    v.flags[dst] = v.flags_in[i];
    v.key[dst] = c;
    v.perm[dst] = mine;
    // the reference permutes its own storage in the sand path (src/neighbors/Neighbors.cpp:296-300):
    // the sorted slot becomes the particle's reference slot.  The fluid path keeps its storage order.
    v.orig[dst] = reset_orig ? dst : mine;
}

int lgpu_launch_reorder(lgpu_ctx* c, bool reset_orig) {
    if (c->n == 0) return LGPU_OK;
    View v = lgpu_make_view(c);
    k_scatter_ids<<<lgpu_blocks(c->n), LGPU_BLOCK, 0, c->stream>>>(v);
    k_reorder<<<lgpu_blocks(c->n), LGPU_BLOCK, 0, c->stream>>>(v, reset_orig ? 1 : 0);
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
    int key = cell_id_checked(g, f3(pos[i]), counters);
    keys[i] = key;
    ranks[i] = atomicAdd(&counts[key], 1);
}
__global__ void k_solid_scatter(const int* keys, const int* ranks, const int* starts, int n, int* tmp) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    tmp[starts[keys[i]] + ranks[i]] = i;
}
__global__ void k_solid_reorder(const float4* pos_in, const int* keys, const int* starts, const int* tmp, int n, float4* pos_out, int* orig_out) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int i = tmp[s];
    int c = keys[i];
    int b = starts[c], e = starts[c + 1];
    int r = 0;
    for (int u = b; u < e; u++) r += (tmp[u] < i) ? 1 : 0;
    pos_out[b + r] = pos_in[i];
    orig_out[b + r] = i;
}

int lgpu_sort_solids(lgpu_ctx* c) {
    const int n = c->n_solid, C = c->g.C;
    CUDA_TRY(cudaMemsetAsync(c->solid_cell_start, 0, sizeof(int) * ((size_t)C + 1), c->stream));
    if (n > 0) {
        // reuse the sand scratch (key_in / rank_in / tmp_id are dead between steps) when it is large
        // enough, otherwise allocate temporaries
        int *keys, *ranks, *tmp, *counts;
        CUDA_TRY(cudaMalloc(&keys, sizeof(int) * n));
        CUDA_TRY(cudaMalloc(&ranks, sizeof(int) * n));
        CUDA_TRY(cudaMalloc(&tmp, sizeof(int) * n));
        CUDA_TRY(cudaMalloc(&counts, sizeof(int) * ((size_t)C + 1)));
        CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(int) * ((size_t)C + 1), c->stream));
        k_solid_keys<<<lgpu_blocks(n), LGPU_BLOCK, 0, c->stream>>>(c->g, c->solid_pos_unsorted, n, keys, ranks, counts, c->counters);
        int st = lgpu_launch_scan_cells(c, counts, c->solid_cell_start, C, false);
        if (st) return st;
        k_solid_scatter<<<lgpu_blocks(n), LGPU_BLOCK, 0, c->stream>>>(keys, ranks, c->solid_cell_start, n, tmp);
        k_solid_reorder<<<lgpu_blocks(n), LGPU_BLOCK, 0, c->stream>>>(c->solid_pos_unsorted, keys, c->solid_cell_start, tmp, n, c->solid_pos, c->solid_orig);
        c->launches += 3;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        cudaFree(keys); cudaFree(ranks); cudaFree(tmp); cudaFree(counts);
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
    CUDA_TRY(cudaMalloc(&tmpctx.scan_block_sums, sizeof(int) * ((size_t)nb + 1)));
    CUDA_TRY(cudaMemcpyAsync(d_keys, keys, sizeof(int) * n, cudaMemcpyHostToDevice, tmpctx.stream));
    CUDA_TRY(cudaMemsetAsync(d_counts, 0, sizeof(int) * ((size_t)num_cells + 1), tmpctx.stream));
    k_cs_hist<<<lgpu_blocks(n), LGPU_BLOCK, 0, tmpctx.stream>>>(d_keys, n, d_ranks, d_counts);
    int st = lgpu_launch_scan_cells(&tmpctx, d_counts, d_starts, num_cells, false);
    if (st) return st;
    k_solid_scatter<<<lgpu_blocks(n), LGPU_BLOCK, 0, tmpctx.stream>>>(d_keys, d_ranks, d_starts, n, d_tmp);
    k_cs_rank<<<lgpu_blocks(n), LGPU_BLOCK, 0, tmpctx.stream>>>(d_keys, d_starts, d_tmp, n, d_sorted);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(sorted, d_sorted, sizeof(int) * n, cudaMemcpyDeviceToHost, tmpctx.stream));
    CUDA_TRY(cudaStreamSynchronize(tmpctx.stream));
    cudaFree(d_keys); cudaFree(d_ranks); cudaFree(d_tmp); cudaFree(d_sorted); cudaFree(d_counts); cudaFree(d_starts);
    cudaFree(tmpctx.scan_block_sums);
    cudaStreamDestroy(tmpctx.stream);
    return LGPU_OK;
}
