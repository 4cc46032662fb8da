// lgpu_neighbors.cu — builds the per-particle neighbour table once per substep.
// Replaces the list construction of find_neighbors_uniform_grid (src/neighbors/Neighbors.cpp:386-448)
// and find_neighbors_uniform_grid_v1 (:306-361).  In the fluid step the same kernel goes on to the first
// density + lambda pass (src/Simulate.cpp:58-88) on the neighbours it has just found: the predicted
// positions are already staged, one stage fill and one kernel launch less per substep.
#include <stdlib.h>

#include "lgpu_fluid.cuh"

// ------------------------------------------------------------------------------------------
// work list: the non-empty bricks of this substep and their descriptors (one warp per brick)
// ------------------------------------------------------------------------------------------
// Full bricks are appended from the front of brick_rec and sparse ones (surface, spray) from the back, so that
// the persistent blocks take the heavy bricks first and the light ones even out the tail.  The descriptor is put
// together in shared memory and written out in one piece; it also reserves the brick's table block.
#define LGPU_DESC_WARPS 8
__global__ void __launch_bounds__(LGPU_DESC_WARPS * 32) k_brick_desc(const __grid_constant__ View v) {
    __shared__ BrickRec recs[LGPU_DESC_WARPS];
    __shared__ int cs32[LGPU_DESC_WARPS][LGPU_HCOLS][LGPU_HB];   // sorted sand slot at every cell boundary of every halo column
    __shared__ int ss32[LGPU_DESC_WARPS][LGPU_HCOLS][LGPU_HB];   // ... sorted solid slot
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int brick = blockIdx.x * LGPU_DESC_WARPS + wib;
    pdl_trigger();
    pdl_wait();  // the cell offsets are the sort's
    if (brick >= v.NB) return;
    const Geom& g = v.g;
    const int by = brick / (v.nbX * v.nbZ);
    const int rem = brick - by * (v.nbX * v.nbZ);
    const int bx = rem / v.nbZ, bz = rem - bx * v.nbZ;
    const int cy0 = by * LGPU_BY, cx0 = bx * LGPU_BX, cz0 = bz * LGPU_BZ;
    // own particles of the whole brick: the inner columns, cells cz0 .. cz0 + BZ - 1
    int own_all = 0;
    if (lane < LGPU_OWN_COLS) {
        const int cy = cy0 + lane / LGPU_BX, cx = cx0 + lane % LGPU_BX;
        if (cy < g.gY && cx < g.gX) {
            const int base = cy * g.gXZ + cx * g.gZ;
            own_all = v.cell_start[base + min(cz0 + LGPU_BZ, g.gZ)] - v.cell_start[base + min(cz0, g.gZ)];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) own_all += __shfl_xor_sync(0xffffffffu, own_all, o);
    if (own_all == 0) return;
    BrickRec& rec = recs[wib];
    BrickDesc& d = rec.d;
    int (*cs)[LGPU_HB] = cs32[wib];
    int (*ss)[LGPU_HB] = ss32[wib];
    // boundaries of the cells cz0-1 .. cz0+BZ of every halo column, clipped to the grid (the cell offsets are linear
    // in the cell id with z fastest, so boundary gZ of a column is the next column's start)
    for (int idx = lane; idx < LGPU_HCOLS * LGPU_HB; idx += 32) {
        const int hc = idx / LGPU_HB, t = idx - hc * LGPU_HB;
        const int cy = cy0 - 1 + hc / LGPU_HX, cx = cx0 - 1 + hc % LGPU_HX;
        int val = 0, sval = 0;
        if (cy >= 0 && cy < g.gY && cx >= 0 && cx < g.gX) {
            const int cell = cy * g.gXZ + cx * g.gZ + min(max(cz0 - 1 + t, 0), g.gZ);
            val = v.cell_start[cell];
            if (v.n_solid) sval = v.solid_cell_start[cell];
        }
        cs[hc][t] = val; ss[hc][t] = sval;
    }
    __syncwarp();
    // A brick whose neighbourhood or table block does not fit a block's shared memory is cut along z into 2, 4, ... parts
    // (dense packings, pile-ups); only a part that still does not fit re-walks the stencil (mode 2).
    const int hi = 32 + lane;
    const bool two = lane < LGPU_HCOLS - 32;
    int parts = 1, nz = LGPU_BZ;
    for (;;) {
        bool fits = true;
        for (int za = 0; za < LGPU_BZ; za += nz) {
            const int zb = min(za + nz, LGPU_BZ);   // part = own cells za .. zb-1 of the brick, halo cells za-1 .. zb
            int slots = cs[lane][zb + 2] - cs[lane][za] + ss[lane][zb + 2] - ss[lane][za];
            if (two) slots += cs[hi][zb + 2] - cs[hi][za] + ss[hi][zb + 2] - ss[hi][za];
            int own = 0;
            if (lane < LGPU_OWN_COLS) {
                const int hc = (lane / LGPU_BX + 1) * LGPU_HX + (lane % LGPU_BX) + 1;
                own = cs[hc][zb + 1] - cs[hc][za + 1];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { slots += __shfl_xor_sync(0xffffffffu, slots, o); own += __shfl_xor_sync(0xffffffffu, own, o); }
            if (slots + LGPU_DUMMY_SLOTS > v.stage_slots || ((own + 1) & ~1) > LGPU_ROW_CAP) fits = false;
        }
        if (fits || nz == 1) break;
        nz = (nz + 1) >> 1; parts++;
    }
    for (int za = 0; za < LGPU_BZ; za += nz) {
        const int zb = min(za + nz, LGPU_BZ);
        // own runs of the part
        int own_len = 0, own_a = 0;
        if (lane < LGPU_OWN_COLS) {
            const int hc = (lane / LGPU_BX + 1) * LGPU_HX + (lane % LGPU_BX) + 1;
            own_a = cs[hc][za + 1];
            own_len = cs[hc][zb + 1] - own_a;
        }
        const int oinc = warp_incl_scan_i(own_len, lane);
        const int n_own = __shfl_sync(0xffffffffu, oinc, LGPU_OWN_COLS - 1);
        if (n_own == 0) continue;
        // lane: halo column `lane`; lanes 0..3 also column 32 + lane.  Stage slots: dummies, sand columns 0..35, solid columns 0..35
        const int len0 = cs[lane][zb + 2] - cs[lane][za];
        const int len1 = two ? cs[hi][zb + 2] - cs[hi][za] : 0;
        const int slen0 = ss[lane][zb + 2] - ss[lane][za];
        const int slen1 = two ? ss[hi][zb + 2] - ss[hi][za] : 0;
        const int inc0 = warp_incl_scan_i(len0, lane), tot0 = __shfl_sync(0xffffffffu, inc0, 31);
        const int inc1 = warp_incl_scan_i(len1, lane), tot1 = __shfl_sync(0xffffffffu, inc1, 31);
        const int sinc0 = warp_incl_scan_i(slen0, lane), stot0 = __shfl_sync(0xffffffffu, sinc0, 31);
        const int sinc1 = warp_incl_scan_i(slen1, lane), stot1 = __shfl_sync(0xffffffffu, sinc1, 31);
        const int solid_base = LGPU_DUMMY_SLOTS + tot0 + tot1;
        const int n_slots = solid_base + stot0 + stot1;
        __syncwarp();  // (the previous part's record has been written out)
        d.col[lane].g0 = cs[lane][za]; d.col[lane].len = len0; d.col[lane].s0 = LGPU_DUMMY_SLOTS + inc0 - len0; d.col[lane].pad = 0;
        d.scol[lane].g0 = ss[lane][za]; d.scol[lane].len = slen0; d.scol[lane].s0 = solid_base + sinc0 - slen0; d.scol[lane].pad = 0;
        if (two) {
            d.col[hi].g0 = cs[hi][za]; d.col[hi].len = len1; d.col[hi].s0 = LGPU_DUMMY_SLOTS + tot0 + inc1 - len1; d.col[hi].pad = 0;
            d.scol[hi].g0 = ss[hi][za]; d.scol[hi].len = slen1; d.scol[hi].s0 = solid_base + stot0 + sinc1 - slen1; d.scol[hi].pad = 0;
        }
        __syncwarp();
        const int n_pad = (n_own + 1) & ~1;
        const int mode = (n_slots > v.stage_slots || n_pad > LGPU_ROW_CAP) ? 2 : 0;
        // boundaries -> stage slots (16 bits; a part whose neighbourhood does not fit never looks at them); the boundaries
        // past the part's last halo cell repeat its end, so that the build's cell search never runs past the part
        for (int idx = lane; idx < LGPU_HCOLS * LGPU_HB; idx += 32) {
            const int hc = idx / LGPU_HB, t = idx - hc * LGPU_HB;
            const int tt = min(za + t, zb + 2);
            rec.cs[hc][t] = (unsigned short)min(d.col[hc].s0 + (cs[hc][tt] - d.col[hc].g0), 65535);
        }
        if (lane < LGPU_OWN_COLS) {
            const int hc = (lane / LGPU_BX + 1) * LGPU_HX + (lane % LGPU_BX) + 1;
            d.own_g0[lane] = own_a;
            d.own_s0[lane] = d.col[hc].s0 + (own_a - d.col[hc].g0);
            d.own_prefix[lane + 1] = oinc;
        }
        int w = 0;
        if (lane == 0) {
            const bool full = n_own >= 256;
            const int k = atomicAdd(&v.brick_ctl[full ? 0 : 1], 1);
            w = full ? k : v.rec_cap - 1 - k;   // record index; the work index of a sparse brick is n_full + k (see rec_of_work)
            d.own_prefix[0] = 0;
            d.brick = brick; d.mode = mode; d.n_own = n_own; d.n_slots = n_slots;
            d.solid_base = solid_base; d.maxg = 0; d.n_pad = n_pad;
            d.tab_off = mode == 0 ? atomicAdd(&v.brick_ctl[2], n_pad * (1 + LGPU_MG)) : 0;
            d.cy0 = cy0; d.cx0 = cx0; d.cz0 = cz0 + za;
            d.work = w;  // (= the record index)
        }
        w = __shfl_sync(0xffffffffu, w, 0);
        __syncwarp();
        int4* gd = reinterpret_cast<int4*>(&v.brick_rec[w]);
        const int4* sd = reinterpret_cast<const int4*>(&rec);
        for (int t = lane; t < (int)(sizeof(BrickRec) / 16); t += 32) gd[t] = sd[t];
    }
}

// ------------------------------------------------------------------------------------------
// table build
// ------------------------------------------------------------------------------------------
// Collects the 16-bit codes of a table row in the lane's row of the slot's table block (shared memory, [group][own
// particle]); the finished block goes to global memory with one bulk copy, and the row is still at hand for the first
// density + lambda pass.
// entry number cnt >= M of a row goes into the list's spill chunk (rare).  Out of line, and everything by value: an
// object whose address is passed to an out-of-line function would have to live in local memory.  state = address of
// the list's spill chunk (0: none yet) | bit 63: the list is lost (no chunk left, or more than M + LGPU_SPILL entries:
// the solver passes re-walk).
#define LGPU_ROW_LOST (1ULL << 63)
__device__ __noinline__ unsigned long long row_put_spill(const View& v, int i, int cnt, uint32_t code, unsigned long long state) {
    const int k = cnt - 4 * LGPU_MG;
    if (k == 0) {
        const int chunk = atomicAdd(&v.brick_ctl[3], 1);
        if (chunk < v.spill_cap) { state = (unsigned long long)(uintptr_t)(v.nbr_spill + (size_t)chunk * (LGPU_SPILL / 4)); v.nbr_ovf[i] = chunk; }
        else state = LGPU_ROW_LOST;
    }
    if (k >= LGPU_SPILL) state |= LGPU_ROW_LOST;
    if (!(state & LGPU_ROW_LOST)) reinterpret_cast<unsigned short*>((uintptr_t)state)[k] = (unsigned short)(code << 4);
    return state;
}
struct RowWriter {
    uint32_t base, stride;    // shared address of the lane's first group; bytes between groups
    int cnt;
    unsigned long long state; // see row_put_spill
    __device__ __forceinline__ void init(uint32_t row_addr, uint32_t row_stride) {
        base = row_addr; stride = row_stride;
        cnt = 0; state = 0;
    }
    __device__ __forceinline__ bool lost() const { return (state & LGPU_ROW_LOST) != 0; }
    // (a stored code is the stage slot x 16: the byte offset the solver passes add to the stage's address)
    __device__ __forceinline__ void put(int k, uint32_t code) {
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(base + (uint32_t)(k >> 2) * stride + (uint32_t)(k & 3) * 2u), "h"((unsigned short)(code << 4)) : "memory");
    }
    __device__ __forceinline__ void emit(const View& v, int i, uint32_t code) {
        if (cnt < 4 * LGPU_MG) put(cnt, code);
        else state = row_put_spill(v, i, cnt, code, state);
        cnt++;
    }
    __device__ __forceinline__ void emit_unchecked(uint32_t code) { put(cnt, code); cnt++; }  // the caller has checked cnt + (codes to come) <= M
    __device__ __forceinline__ void finish(uint32_t pad) {  // completes the last group with the padding code
        if (cnt <= 4 * LGPU_MG) { for (int k = cnt; k & 3; k++) put(k, pad); }
        else if (!lost()) {
            unsigned short* spill = reinterpret_cast<unsigned short*>((uintptr_t)state);
            for (int k = cnt - 4 * LGPU_MG; k & 3; k++) spill[k] = (unsigned short)(pad << 4);
        }
    }
};

// Out-of-line path of the table build: particles with solids in their neighbourhood or a very dense column walk the
// stencil over the global storage in the reference's nested order (per cell: sand then solids in the fluid lists,
// solids then sand in the sand lists) and translate what they find into stage slots of their brick.
template <bool SAND>
__device__ __noinline__ int build_row_walk(const View& v, const BrickDesc& d, int i, uint32_t row_addr, uint32_t row_stride, F3 xi, int ly, int lx, uint32_t pad) {
    RowWriter w;
    w.init(row_addr, row_stride);
    walk<SAND>(v, i, xi, [&](int j, int r) {
        const int hc = (ly + r / 3) * LGPU_HX + lx + r % 3;  // halo column of stencil column r = 3 (dy + 1) + (dx + 1)
        if (j >= 0) w.emit(v, i, (uint32_t)(d.col[hc].s0 + (j - d.col[hc].g0)));
        else w.emit(v, i, (uint32_t)(d.scol[hc].s0 + (~j - d.scol[hc].g0)));
    });
    w.finish(pad);
    return w.lost() ? -w.cnt : w.cnt;
}

template <bool SAND, int LM>
__global__ void __launch_bounds__(LGPU_BRICK_THREADS, LGPU_CTAS_PER_SM) k_build_table(const __grid_constant__ View v, const __grid_constant__ FluidParams fp, int* cursor) {
    extern __shared__ unsigned char smem_raw[];
    brick_loop<true>(v, v.x0, cursor, smem_raw, [&](const Chunk& ck) -> int {
        const Geom& g = v.g;
        const BrickDesc& d = *ck.d;
        const int i = ck.i, slot = ck.slot, q = ck.q;
        const uint32_t stage_addr = ck.stage_addr;
        const float4 x0i = d.mode == 0 ? lds128(slot_addr(stage_addr, (uint32_t)slot)) : v.x0[i];
        const F3 xi = f3(x0i);
        int word, ng = 0;
        const int flags = g.slab ? v.flags[i] : 0;
        // Ghost of a neighbouring slab: read by others, its position never updated here.  With a two-column ghost layer
        // the ghosts of the column next to the owned ones still get a list: all their neighbours are present, so their
        // lambda is computed locally instead of being sent by the owner after every lambda pass.
        int ghost_word = 0;
        if (flags & LGPU_FLAG_GHOST) {
            ghost_word = LGPU_CNT_GHOST;
            if (g.gw == 2 && !SAND) {
                const int gx = __float2int_rz(__fdiv_rn(xi.x, g.cell_size));  // global cell column (get_cell_id's arithmetic)
                if (gx == g.x_lo - 1 || gx == g.x_hi) ghost_word = LGPU_CNT_GHOST_INNER;
            }
        }
        if (ghost_word == LGPU_CNT_GHOST) {
            word = LGPU_CNT_GHOST;
        } else if (d.mode != 0) {
            // neighbourhood larger than a ring slot: count only, the solver passes re-walk the stencil
            int cnt = 0;
            walk<SAND>(v, i, f3(v.x0[i]), [&](int, int) { cnt++; });
            word = cnt | LGPU_CNT_WALK | ghost_word;
            atomicAdd(&v.counters[1], 1ULL);
        } else {
            const int ly = q / LGPU_BX, lx = q % LGPU_BX;
            const int hown = (ly + 1) * LGPU_HX + lx + 1;
            // the particle's cell along its own run: boundary index tz with cs[hown][tz] <= slot < cs[hown][tz + 1]
            int tz = 1;
#pragma unroll
            for (int t = 2; t <= LGPU_BZ; t++) tz += slot >= ck.cs[hown][t] ? 1 : 0;
            // per stencil column: the candidates are the stage slots [cb, ce) — the three cells z-1..z+1 of a column are
            // contiguous in the sorted storage and therefore in the stage
            int cb[9], ce[9];
            bool slow = false;  // solids in the neighbouring columns
#pragma unroll
            for (int r = 0; r < 9; r++) {
                const int hc = (ly + r / 3) * LGPU_HX + lx + r % 3;
                cb[r] = ck.cs[hc][tz - 1]; ce[r] = ck.cs[hc][tz + 2];
                if (d.scol[hc].len > 0) slow = true;  // (the solid range of a halo column spans the brick's z extent)
            }
            const uint32_t self_code = (uint32_t)slot;
            // padding: sand = a far-away dummy (no contact); fluid = the particle itself (zero separation: every term of
            // the branch-free fluid bodies vanishes)
            const uint32_t pad = SAND ? 0u : self_code;
            int cnt;
            if (slow) {
                cnt = build_row_walk<SAND>(v, d, i, ck.row_addr, ck.row_stride, xi, ly, lx, pad);
            } else {
                RowWriter w;
                w.init(ck.row_addr, ck.row_stride);
                const unsigned long long xi_xy = pack2(xi.x, xi.y);
                // No solid near: the reference order is simply ascending sorted slot over the 9 columns (fluid: self
                // included; sand: self skipped — SURVEY F7).  Per column, 32 candidates at a time (one round unless the
                // cells are crowded): test them with the Exact predicate into a hit mask, then emit the hits.
#pragma unroll
                for (int r = 0; r < 9; r++) {
                    for (int c0 = cb[r]; c0 < ce[r]; c0 += 32) {
                        const int n = min(ce[r] - c0, 32);
                        const uint32_t first = (uint32_t)c0;
                        // `out` collects one bit per candidate, shifted in from the right: candidate t ends up at bit n-1-t.
                        // The bit is the sign of h2 - r2 (set <=> r2 > h2; the rounded difference has the exact sign and is
                        // +0 on equality), so a test costs the separately rounded operations of the reference's
                        // predicate (src/neighbors/Neighbors.cpp:433-435; x and y as a packed pair: dist2_exact) plus one FADD and one funnel shift.
                        uint32_t out = 0;
                        const uint32_t a = slot_addr(stage_addr, first);
#pragma unroll 4
                        for (int t = 0; t < n; t++) {
                            const float4 pj = lds128(a + 16u * (uint32_t)t);
                            out = __funnelshift_l(__float_as_uint(__fsub_rn(g.h2, dist2_exact(xi_xy, xi.z, pj))), out, 1);
                        }
                        uint32_t m = ~out & (0xffffffffu >> (32 - n));  // hits; candidate t at bit n-1-t
                        const uint32_t top = first + (uint32_t)(n - 1);  // code of bit k = top - k
                        if (SAND && r == 4 && self_code >= first && self_code <= top) m &= ~(1u << (top - self_code));
                        if (w.cnt + __popc(m) <= 4 * LGPU_MG) {
                            while (m) {  // ascending candidate = descending bit
                                const uint32_t k = 31u - (uint32_t)__clz(m);
                                m ^= 1u << k;
                                w.emit_unchecked(top - k);
                            }
                        } else {
                            while (m) {
                                const uint32_t k = 31u - (uint32_t)__clz(m);
                                m ^= 1u << k;
                                w.emit(v, i, top - k);
                            }
                        }
                    }
                }
                w.finish(pad);
                cnt = w.lost() ? -w.cnt : w.cnt;
            }
            // (cnt < 0: the list did not fit the table row plus a spill chunk)
            word = cnt < 0 ? -cnt : cnt;
            if (cnt < 0 || (v.M < 4 * LGPU_MG && cnt > v.M)) { word |= LGPU_CNT_WALK; atomicAdd(&v.counters[1], 1ULL); }
            else ng = (min(cnt, 4 * LGPU_MG) + 3) >> 2;
            word |= ghost_word;  // (LGPU_CNT_GHOST_INNER or nothing)
        }
        v.nbr_cnt[i] = word;
        const int mword = (word & LGPU_CNT_WALK) ? (word & ~LGPU_CNT_MASK) : word;  // (a re-walked row's length is not needed)
        if (d.mode == 0) *ck.meta = make_uint2((uint32_t)i, (uint32_t)(mword | (slot << LGPU_CNT_SLOT_SHIFT)));
        if (!SAND && LM != LM_NONE) {
            // first density + lambda pass on the list just collected
            if (!(word & LGPU_CNT_GHOST)) fluid_lambda_particle<LM>(v, fp, ck, v.x0, mword, xi);
        }
        return ng;
    });
}

static int brick_grid(const lgpu_ctx* c) { return LGPU_CTAS_PER_SM * c->num_sms; }

template <bool SAND, int LM>
static int launch_build(lgpu_ctx* c, const View& v, const FluidParams& fp, bool pdl) {
    CUDA_TRY(launch_pdl(k_build_table<SAND, LM>, brick_grid(c), LGPU_BRICK_THREADS, LGPU_BRICK_SMEM_BUILD, c->stream, pdl, v, fp, c->brick_ctl + 8 + c->pass));
    return LGPU_OK;
}

// lambda_mode: LM_* for the fluid step (the build kernel also runs the first density + lambda pass), LM_NONE otherwise
int lgpu_launch_build_table(lgpu_ctx* c, bool sand_order, const lgpu_step_params& p, int lambda_mode) {
    if (c->n == 0) return LGPU_OK;  // (slab mode: the solver drivers still run the refresh protocol)
    View v = lgpu_make_view(c);
    FluidParams fp = lgpu_make_fluid_params(c->g, p);
    const bool pdl = lgpu_pdl_enabled(c);
    // (slab mode: the predecessor of k_brick_desc is a plain launch of the refresh-list kernels, not a kernel of this chain)
    CUDA_TRY(launch_pdl(k_brick_desc, (c->NB + LGPU_DESC_WARPS - 1) / LGPU_DESC_WARPS, LGPU_DESC_WARPS * 32, 0, c->stream, pdl && !c->g.slab, v));
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
    const int n_full = v.brick_ctl[0], n_work = n_full + v.brick_ctl[1];
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const BrickDesc& d = v.brick_rec[rec_of_work(v, w, n_full)].d;
        const uint2* tab = v.nbr16 + d.tab_off;
        for (int q = 0; q < LGPU_OWN_COLS; q++) {
            const int len = d.own_prefix[q + 1] - d.own_prefix[q];
            for (int t = threadIdx.x; t < len; t += blockDim.x) {
                const int i = d.own_g0[q] + t, p = d.own_prefix[q] + t;
                if (v.g.slab ? (v.flags[i] & LGPU_FLAG_GHOST) != 0 : false) continue;
                long o = offsets[i];
                const int word = v.nbr_cnt[i];
                if (!(word & LGPU_CNT_WALK) && d.mode == 0) {
                    const int cnt = word & LGPU_CNT_MASK;
                    for (int k = 0; k < cnt; k++) {
                        const uint2 g4 = k < 4 * LGPU_MG ? tab[(size_t)(1 + (k >> 2)) * d.n_pad + p]
                                                         : v.nbr_spill[(size_t)v.nbr_ovf[i] * (LGPU_SPILL / 4) + ((k - 4 * LGPU_MG) >> 2)];
                        const uint32_t pair = (k & 2) ? g4.y : g4.x;
                        const int code = (int)((k & 1) ? pair >> 16 : pair & 0xffffu);
                        const int j = decode_code(d, code);
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
    CUDA_TRY(cudaFuncSetAttribute(k_build_table<SAND, LM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LGPU_BRICK_SMEM_BUILD));
    return LGPU_OK;
}
// Loads every kernel of this file on the CURRENT device and opts the staged ones into their dynamic
// shared memory.  cudaFuncSetAttribute is per device: lgpu_create calls this for every context.
int lgpu_preload_neighbors() {
    int st = preload_build<true, LM_NONE>() | preload_build<false, LM_NONE>() | preload_build<false, LM_FAST>() |
             preload_build<false, LM_EXACT>() | preload_build<false, LM_POLY6>() | preload_build<false, LM_GENERIC>();
    LGPU_PRELOAD(k_brick_desc); LGPU_PRELOAD(k_dump_nbr<true>); LGPU_PRELOAD(k_dump_nbr<false>);
    return st ? LGPU_ERR_CUDA : LGPU_OK;
}
