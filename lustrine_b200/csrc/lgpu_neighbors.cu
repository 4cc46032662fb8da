// lgpu_neighbors.cu — builds the per-particle neighbour table once per substep.
// Replaces the list construction of find_neighbors_uniform_grid (src/neighbors/Neighbors.cpp:386-448)
// and find_neighbors_uniform_grid_v1 (:306-361).  In the fluid step the same kernel goes on to the first
// density + lambda pass (src/Simulate.cpp:58-88) on the neighbours it has just found: the predicted
// positions are already staged, one stage fill and one kernel launch less per substep.
#include <stdlib.h>

#include "lgpu_fluid.cuh"

// ------------------------------------------------------------------------------------------
// work list: the non-empty bricks of this substep (one thread per brick)
// ------------------------------------------------------------------------------------------
// Full bricks are appended from the front of brick_work and sparse ones (surface, spray) from the back, so that
// the persistent blocks take the heavy bricks first and the light ones even out the tail.
__global__ void __launch_bounds__(128) k_brick_list(View v) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= v.NB) return;
    const Geom& g = v.g;
    const int by = b / (v.nbX * v.nbZ);
    const int rem = b - by * (v.nbX * v.nbZ);
    const int bx = rem / v.nbZ, bz = rem - bx * v.nbZ;
    const int z0 = bz * LGPU_BZ, z1 = min(z0 + LGPU_BZ, g.gZ);
    int n = 0;
#pragma unroll
    for (int q = 0; q < LGPU_OWN_COLS; q++) {
        const int cy = by * LGPU_BY + q / LGPU_BX, cx = bx * LGPU_BX + q % LGPU_BX;
        if (cy < g.gY && cx < g.gX) {
            const int base = cy * g.gXZ + cx * g.gZ;
            n += v.cell_start[base + z1] - v.cell_start[base + z0];
        }
    }
    if (n == 0) return;
    // (warp-aggregated by the compiler; the order inside the two lists does not affect any result)
    if (n >= 256) v.brick_work[atomicAdd(&v.brick_ctl[0], 1)] = b;
    else v.brick_work[v.NB - 1 - atomicAdd(&v.brick_ctl[1], 1)] = b;
}

// ------------------------------------------------------------------------------------------
// table build
// ------------------------------------------------------------------------------------------
// Appends 16-bit codes to a table row.  Four consecutive codes share one 8-byte group and the groups
// are strided by the capacity (nbr16 layout in lgpu_internal.cuh): the codes are packed in a 64-bit
// shift register and every fourth emit stores one group.
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

// Out-of-line path of the table build: particles with solids in their neighbourhood or a very dense column walk the
// stencil over the global storage in the reference's nested order (per cell: sand then solids in the fluid lists,
// solids then sand in the sand lists) and translate what they find into stage slots of their brick.
template <bool SAND>
__device__ __noinline__ int2 build_row_walk(const View& v, const BrickInfo& info, int i, F3 xi, int ly, int lx, uint32_t pad) {
    RowWriter w;
    w.init(v, i);
    walk<SAND>(v, i, xi, [&](int j, int r) {
        const int hc = (ly + r / 3) * LGPU_HX + lx + r % 3;  // halo column of stencil column r = 3 (dy + 1) + (dx + 1)
        if (j >= 0) w.emit((uint32_t)(info.col_s0[hc] + (j - info.col_g0[hc])));
        else w.emit((uint32_t)(info.scol_s0[hc] + (~j - info.scol_g0[hc])));
    });
    w.finish(pad);
    return make_int2(w.cnt, w.bad ? 1 : 0);
}

template <bool SAND, int LM>
__global__ void __launch_bounds__(LGPU_BRICK_THREADS, 2) k_build_table(const __grid_constant__ View v, const __grid_constant__ FluidParams fp, int* cursor) {
    extern __shared__ unsigned char smem_raw[];
    brick_loop<true>(v, v.x0, cursor, smem_raw, [&](const BrickInfo& info, const float4* stage, int q, int i, int slot) {
        const Geom& g = v.g;
        const uint32_t stage_addr = smem_u32(stage);
        const float4 x0i = info.mode == 0 ? lds128(slot_addr(stage_addr, (uint32_t)slot)) : v.x0[i];
        const F3 xi = f3(x0i);
        int word;
        const int flags = g.slab ? v.flags[i] : 0;
        if (flags & LGPU_FLAG_GHOST) {  // ghost of a neighbouring slab: read by others, never updated here
            word = LGPU_CNT_GHOST;
        } else if (info.mode != 0) {
            // neighbourhood larger than the stage: count only, the solver passes re-walk the stencil
            int cnt = 0;
            walk<SAND>(v, i, f3(v.x0[i]), [&](int, int) { cnt++; });
            word = cnt | LGPU_CNT_WALK;
            atomicAdd(&v.counters[1], 1ULL);
        } else {
            const int ly = q / LGPU_BX, lx = q % LGPU_BX;
            const int hown = (ly + 1) * LGPU_HX + lx + 1;
            // the particle's cell along its own run: boundary index tz with cs[hown][tz] <= slot < cs[hown][tz + 1]
            int tz = 1;
#pragma unroll
            for (int t = 2; t <= LGPU_BZ; t++) tz += slot >= info.cs[hown][t] ? 1 : 0;
            // per stencil column: the candidates are the stage slots [cb, ce) — the three cells z-1..z+1 of a column are
            // contiguous in the sorted storage and therefore in the stage
            int cb[9], ce[9];
            bool slow = false;  // solids in the neighbouring columns, or a column with more than 32 candidates
#pragma unroll
            for (int r = 0; r < 9; r++) {
                const int hc = (ly + r / 3) * LGPU_HX + lx + r % 3;
                cb[r] = info.cs[hc][tz - 1]; ce[r] = info.cs[hc][tz + 2];
                if (ce[r] - cb[r] > 32) slow = true;
            }
            if (v.n_solid) {
                // (one compare per column: the solid range of a halo column spans the brick's z extent)
#pragma unroll
                for (int r = 0; r < 9; r++) {
                    const int hc = (ly + r / 3) * LGPU_HX + lx + r % 3;
                    if (info.scol_len[hc] > 0) slow = true;
                }
            }
            const uint32_t self_code = (uint32_t)slot;
            // padding: sand = a far-away dummy (no contact); fluid = the particle itself (zero separation: every term of
            // the branch-free fluid bodies vanishes)
            const uint32_t pad = SAND ? 0u : self_code;
            int cnt;
            bool bad;
            if (slow) {
                const int2 r = build_row_walk<SAND>(v, info, i, xi, ly, lx, pad);
                cnt = r.x; bad = r.y != 0;
            } else {
                RowWriter w;
                w.init(v, i);
                // No solid near: the reference order is simply ascending sorted slot over the 9 columns (fluid: self
                // included; sand: self skipped — SURVEY F7).  Per column: test the candidates with the Exact predicate
                // into a hit mask, then emit the hits.
#pragma unroll
                for (int r = 0; r < 9; r++) {
                    const int n = ce[r] - cb[r];
                    if (n == 0) continue;
                    const uint32_t first = (uint32_t)cb[r];
                    // `out` collects one bit per candidate, shifted in from the right: candidate t ends up at bit n-1-t.
                    // The bit is the sign of h2 - r2 (set <=> r2 > h2; the rounded difference has the exact sign and is
                    // +0 on equality), so a test costs the 8 separately rounded operations of the reference's
                    // predicate (src/neighbors/Neighbors.cpp:433-435) plus one FADD and one funnel shift.
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
            word = cnt;
            if (cnt > v.M || bad) { word |= LGPU_CNT_WALK; atomicAdd(&v.counters[1], 1ULL); }
        }
        v.nbr_cnt[i] = word;
        if (!SAND && LM != LM_NONE) {
            // first density + lambda pass on the list just written (the thread re-reads its own row)
            if (!(word & LGPU_CNT_GHOST)) fluid_lambda_particle<LM>(v, fp, info, stage, v.x0, i, word, xi);
        }
    });
}

static int brick_grid(const lgpu_ctx* c) { return 2 * c->num_sms; }

template <bool SAND, int LM>
static int launch_build(lgpu_ctx* c, const View& v, const FluidParams& fp, bool pdl) {
    CUDA_TRY(launch_pdl(k_build_table<SAND, LM>, brick_grid(c), LGPU_BRICK_THREADS, LGPU_BRICK_SMEM, c->stream, pdl, v, fp, c->brick_ctl + 8 + c->pass));
    return LGPU_OK;
}

// lambda_mode: LM_* for the fluid step (the build kernel also runs the first density + lambda pass), LM_NONE otherwise
int lgpu_launch_build_table(lgpu_ctx* c, bool sand_order, const lgpu_step_params& p, int lambda_mode) {
    if (c->n == 0) return LGPU_OK;  // (slab mode: the solver drivers still run the refresh protocol)
    View v = lgpu_make_view(c);
    FluidParams fp = lgpu_make_fluid_params(c->g, p);
    k_brick_list<<<(c->NB + 127) / 128, 128, 0, c->stream>>>(v);
    static const bool pdl_env = !(getenv("LGPU_PDL") && atoi(getenv("LGPU_PDL")) == 0);
    const bool pdl = pdl_env && !c->phase_timing && !c->use_graph;
    int st;
    if (sand_order) st = launch_build<true, LM_NONE>(c, v, fp, pdl);
    else switch (lambda_mode) {
        case LM_FAST: st = launch_build<false, LM_FAST>(c, v, fp, pdl); break;
        case LM_EXACT: st = launch_build<false, LM_EXACT>(c, v, fp, pdl); break;
        case LM_POLY6: st = launch_build<false, LM_POLY6>(c, v, fp, pdl); break;
        case LM_GENERIC: st = launch_build<false, LM_GENERIC>(c, v, fp, pdl); break;
        default: st = launch_build<false, LM_NONE>(c, v, fp, pdl); break;
    }
    if (st) return st;
    c->pass++;
    c->launches += 2;
    CUDA_TRY(cudaGetLastError());
    return LGPU_OK;
}

// Neighbour lists as the solver passes see them (tests, lgpu_dump): one block per brick of the work list; codes are
// translated back to sorted slots through the brick's descriptor, rows the table does not hold are re-walked.
template <bool SAND>
__global__ void __launch_bounds__(128) k_dump_nbr(View v, const long* __restrict__ offsets, int* __restrict__ flat) {
    __shared__ BrickInfo info;
    const int n_full = v.brick_ctl[0], n_work = n_full + v.brick_ctl[1];
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        __syncthreads();
        const BrickDesc& d = v.brick_desc[w];
        if (threadIdx.x == 0) {
            int s = LGPU_DUMMY_SLOTS;
            for (int hc = 0; hc < LGPU_HCOLS; hc++) { info.col_g0[hc] = d.g0[hc]; info.col_s0[hc] = s; info.col_len[hc] = d.len[hc]; s += d.len[hc]; }
            for (int hc = 0; hc < LGPU_HCOLS; hc++) { info.scol_g0[hc] = d.sg0[hc]; info.scol_s0[hc] = s; info.scol_len[hc] = d.slen[hc]; s += d.slen[hc]; }
        }
        __syncthreads();
        for (int q = 0; q < LGPU_OWN_COLS; q++) {
            for (int t = threadIdx.x; t < d.own_len[q]; t += blockDim.x) {
                const int i = d.own_g0[q] + t;
                if (v.g.slab ? (v.flags[i] & LGPU_FLAG_GHOST) != 0 : false) continue;
                long o = offsets[i];
                const int word = v.nbr_cnt[i];
                if (!(word & LGPU_CNT_WALK)) {
                    const int cnt = word & LGPU_CNT_MASK;
                    for (int k = 0; k < cnt; k++) {
                        const uint2 g4 = v.nbr16[(size_t)(k >> 2) * v.cap + i];
                        const uint32_t pair = (k & 2) ? g4.y : g4.x;
                        const int code = (int)((k & 1) ? pair >> 16 : pair & 0xffffu);
                        const int j = decode_code(info, code);
                        flat[o++] = j >= 0 ? j : v.n + v.solid_orig[~j];
                    }
                } else {
                    walk<SAND>(v, i, f3(v.x0[i]), [&](int j, int) { flat[o++] = j >= 0 ? j : v.n + v.solid_orig[~j]; });
                }
            }
        }
    }
}
int lgpu_launch_dump_nbr(lgpu_ctx* c, bool sand, const long* d_off, int* d_flat) {
    View v = lgpu_make_view(c);
    if (sand) k_dump_nbr<true><<<2 * c->num_sms, 128, 0, c->stream>>>(v, d_off, d_flat);
    else k_dump_nbr<false><<<2 * c->num_sms, 128, 0, c->stream>>>(v, d_off, d_flat);
    CUDA_TRY(cudaGetLastError());
    return LGPU_OK;
}

#define LGPU_PRELOAD(f) do { cudaFuncAttributes a; CUDA_TRY(cudaFuncGetAttributes(&a, f)); } while (0)
template <bool SAND, int LM> static int preload_build() {
    CUDA_TRY(cudaFuncSetAttribute(k_build_table<SAND, LM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LGPU_BRICK_SMEM));
    return LGPU_OK;
}
// Loads every kernel of this file on the CURRENT device and opts the staged ones into their dynamic
// shared memory.  cudaFuncSetAttribute is per device: lgpu_create calls this for every context.
int lgpu_preload_neighbors() {
    int st = preload_build<true, LM_NONE>() | preload_build<false, LM_NONE>() | preload_build<false, LM_FAST>() |
             preload_build<false, LM_EXACT>() | preload_build<false, LM_POLY6>() | preload_build<false, LM_GENERIC>();
    LGPU_PRELOAD(k_brick_list); LGPU_PRELOAD(k_dump_nbr<true>); LGPU_PRELOAD(k_dump_nbr<false>);
    return st ? LGPU_ERR_CUDA : LGPU_OK;
}
