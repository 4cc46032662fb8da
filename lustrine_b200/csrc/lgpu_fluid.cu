// lgpu_fluid.cu — position-based-fluids solver iterations.
// Replaces the two solver loops of Lustrine::simulate_fluid, src/Simulate.cpp:58-88 (density +
// lambda) and :90-113 (delta-p, box collision, velocity/position commit), with s_coor (:7-9),
// resolve_collision (:13-24), cubic_kernel / cubic_kernel_grad (src/Kernels.cpp:6-41) inlined.
//
// Two fused kernels per solver iteration:
//   k_fluid_lambda   reads x* of the neighbours, writes rho_i and lambda_i          (16 B/particle)
//   k_fluid_deltap   reads x*, lambda of the neighbours, writes the corrected x*    (28 B/particle)
//                    and on the last iteration also v and x                         (+24 B/particle)
// The delta-p output is double-buffered (Jacobi); the reference's in-place loop is sequential
// Gauss-Seidel in index order and is compared through the Jacobi oracle.
#include <stdlib.h>

#include "lgpu_neighbors.cuh"

struct FluidParams {
    float dt, rest_density, mass, eps;
    float s_corr_k, s_corr_n;
    float W_dq;       // W(s_corr_dq), hoisted: same value for every pair
    float W_zero;     // W(0)
    float neg_mr;     // -(mass / rest_density)
    // Fast-policy constants
    float c_q;        // kernelFactor / h
    float l_h2;       // cubic_l / (h*h)
    float l_kfh;      // cubic_l / (kernelFactor * h)
    float inv_W_dq, inv_rho0, inv_dt;
    float gA, gB;     // Fast lambda pass: -(m/rho0) * gradW coefficient = gA*q + gB   (q <= 0.5)
    float cA, cB;     // Fast delta-p pass: gradW coefficient = cA*q + cB             (q <= 0.5)
    float kx;         // cubic_k / W(s_corr_dq)
    float mk;         // mass * cubic_k
    // branch-free inner evaluation of the fast kernels (k_fluid_*_fast), in terms of len = |d| and r2 = len^2:
    float fA, fB;     // W/cubic_k = 1 + r2 * (len*fA - fB)                    (6c^3, 6c^2, c = kernelFactor/h)
    float thr2;       // r2 > thr2 <=> q > 0.5: the pair is corrected out of line
    float fgA;        // lambda pass:  -(m/rho0) * gradW coefficient = len*fgA + gB
    float fcA;        // delta-p pass: gradW coefficient = len*fcA + cB
    float xA, xB;     // delta-p pass: W/W(s_corr_dq) = kx + r2 * (len*xA - xB)
    int literal_lambda_index;
};

__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

template <bool POLY6, class P> __device__ __forceinline__ float W_of(const Geom& g, float r) {
    if (POLY6) return poly6_W(g, r);
    return cubic_W<P>(g, r);
}
template <bool POLY6, class P> __device__ __forceinline__ F3 gradW_of(const Geom& g, F3 d) {
    if (POLY6) return spiky_gradW(g, d);
    return cubic_gradW<P>(g, d);
}

// Fast-policy pair evaluation of the cubic spline: W(|d|) and the scalar c with gradW(d) = c * d.
// Algebraically identical to src/Kernels.cpp:6-41: for q <= 0.5, l*q*(3q-2) * d/(rl*h) with
// q = rl/h collapses to (l/h^2)*(3q-2) * d.
__device__ __forceinline__ void cubic_pair_fast(const Geom& g, const FluidParams& fp, float r2, float& Wv, float& coef) {
    float len = sqrt_approx(r2);
    float q = len * fp.c_q;
    Wv = 0.0f;
    coef = 0.0f;
    if (q <= 0.5f) {
        float q2 = q * q;
        Wv = g.cubic_k * (q2 * (6.0f * q - 6.0f) + 1.0f);
        if (len * g.kernel_factor > 1.0e-5f) coef = fp.l_h2 * (3.0f * q - 2.0f);
    } else if (q <= 1.0f) {
        float f = 1.0f - q;
        Wv = g.cubic_k * (2.0f * f * f * f);
        coef = -fp.l_kfh * f * f * rsqrtf(r2);
    }
}

// Fast-policy pair evaluation used by the solver passes: wp = W(|d|)/cubic_k and cf = A*q + B on
// the inner branch q <= 0.5 (the only one list neighbours reach at build time); the
// outer branch and the cut-off are handled out of line.  A, B = the pass's pre-scaled gradient
// constants, outer = its scale of the outer-branch gradient (-l/(kf*h) times the same factor).
__device__ __forceinline__ void cubic_pair_inner(const FluidParams& fp, float r2, float A, float B, float outer, float& wp, float& cf) {
    const float len = sqrt_approx(r2);
    const float q = len * fp.c_q;
    const float t = fmaf(q, 6.0f, -6.0f);
    wp = fmaf(q * q, t, 1.0f);
    cf = fmaf(q, A, B);
    if (q > 0.5f) {
        wp = 0.0f; cf = 0.0f;
        if (q <= 1.0f) {
            const float f = 1.0f - q;
            wp = 2.0f * f * f * f;
            cf = outer * f * f * rsqrtf(r2);
        }
    }
    cf = r2 > 4.0e-10f ? cf : 0.0f;  // rl = |d|*kernelFactor > 1e-5 (src/Kernels.cpp:32)
}

// resolve_collision, src/Simulate.cpp:13-24 (returns 0.01, not min)
__device__ __forceinline__ float resolve_collision(float value, float lo, float hi) {
    if (value <= lo) return 0.01f;
    if (value > hi) return __fsub_rn(hi, 0.01f);
    return value;
}

// ---- density + lambda: src/Simulate.cpp:58-88 ----
template <class P, bool POLY6>
struct LambdaAcc {
    float rho, sum;
    F3 gi;
    __device__ __forceinline__ void init() { rho = 0.0f; sum = 0.0f; gi = f3(0.0f, 0.0f, 0.0f); }
    __device__ __forceinline__ void pair(const Geom& g, const FluidParams& fp, F3 xi, F3 xj) {
        if (P::exact || POLY6) {
            F3 d = vsub<P>(xi, xj);
            float len = vlen<P>(d);
            rho = P::add(rho, P::mul(fp.mass, W_of<POLY6, P>(g, len)));           // :62-64
            F3 gr = vscale<P>(gradW_of<POLY6, P>(g, d), fp.neg_mr);                // :76
            sum = P::add(sum, vdot<P>(gr, gr));                                    // :77
            gi = vsub<P>(gi, gr);                                                  // :78
        } else {
            const float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
            const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            float wp, gs;
            cubic_pair_inner(fp, r2, fp.gA, fp.gB, -fp.neg_mr * fp.l_kfh, wp, gs);
            rho += wp;  // scaled by mass * cubic_k in finish()
            sum = fmaf(gs * gs, r2, sum);
            gi.x = fmaf(-gs, dx, gi.x); gi.y = fmaf(-gs, dy, gi.y); gi.z = fmaf(-gs, dz, gi.z);
        }
    }
    __device__ __forceinline__ float finish(const FluidParams& fp) {
        float lam = 0.0f;
        if (P::exact || POLY6) {
            rho = P::add(rho, P::mul(fp.mass, fp.W_zero));                             // :66
            float Ci = P::sub(P::div(rho, fp.rest_density), 1.0f);                     // :69
            sum = P::add(sum, vdot<P>(gi, gi));                                        // :81
            if (sum > 0.0f) lam = P::div(-Ci, P::add(sum, fp.eps));                    // :83-86
        } else {
            rho = fmaf(rho, fp.mk, fp.mass * fp.W_zero);
            float Ci = rho * fp.inv_rho0 - 1.0f;
            sum += gi.x * gi.x + gi.y * gi.y + gi.z * gi.z;
            if (sum > 0.0f) lam = __fdividef(-Ci, sum + fp.eps);
        }
        return lam;
    }
};

// Reads x* of the neighbours from the stage, writes rho_i, lambda_i and also lambda_i into the w
// lane of the particle's own x* so that the delta-p pass gets (x*_j, lambda_j) in one LDS.128.
template <class P, bool POLY6, bool SOLIDS>
__global__ void __launch_bounds__(LGPU_TILE) k_fluid_lambda(View v, FluidParams fp, float4* __restrict__ cur) {
    extern __shared__ float4 stage[];
    __shared__ BlkDesc d;
    __shared__ uint64_t bar;
    const int i = blockIdx.x * LGPU_TILE + threadIdx.x;
    stage_begin(v, cur, d, &bar, stage);
    const int word = i < v.n ? v.nbr_cnt[i] : LGPU_CNT_GHOST;
    if (word & (LGPU_CNT_GHOST | LGPU_CNT_WALK)) stage_wait(&bar);  // (the table path waits after its row loads)
    if (!(word & LGPU_CNT_GHOST)) {
    const Geom& g = v.g;
    const F3 xi = f3(cur[i]);
    LambdaAcc<P, POLY6> acc;
    acc.init();
    if (!(word & LGPU_CNT_WALK)) {
        replay_neighbors<SOLIDS, false>(v, d, &bar, stage, cur, i, word & LGPU_CNT_MASK,
                                                   [&](float4 pj, uint32_t, int) { acc.pair(g, fp, xi, f3(pj)); });
    } else {
        walk<false>(v, i, f3(v.x0[i]), [&](int j, int) { acc.pair(g, fp, xi, j >= 0 ? f3(cur[j]) : f3(v.solid_pos[~j])); });
    }
    const float lam = acc.finish(fp);
    v.density[i] = acc.rho;
    v.lambda[i] = lam;
    reinterpret_cast<float*>(cur + i)[3] = lam;
    int o = v.orig[i];
    if (o < LGPU_LAMBDA_HEAD) v.lambda_head[o] = lam;  // lambdas[] in reference slot order, for F4
    }
}

// ---- delta-p + box collision (+ commit): src/Simulate.cpp:90-113 ----
template <class P, bool POLY6>
__device__ __forceinline__ void deltap_pair(const Geom& g, const FluidParams& fp, F3 xi, F3 xj, float li, float lj, F3& f) {
    if (P::exact || POLY6) {
        F3 d = vsub<P>(xi, xj);
        float len = vlen<P>(d);
        float x = P::div(W_of<POLY6, P>(g, len), fp.W_dq);
        float sc = P::mul(-fp.s_corr_k, powf_like_libm(x, fp.s_corr_n));     // :7-9
        float w = P::add(P::add(li, lj), sc);
        f = vadd<P>(f, vscale<P>(gradW_of<POLY6, P>(g, d), w));               // :97
    } else {
        const float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
        const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        float wp, cf;
        cubic_pair_inner(fp, r2, fp.cA, fp.cB, -fp.l_kfh, wp, cf);
        const float x = wp * fp.kx;
        const float x2 = x * x;
        const float pw = fp.s_corr_n == 4.0f ? x2 * x2 : __powf(x, fp.s_corr_n);
        const float w = fmaf(-fp.s_corr_k, pw, li + lj) * cf;
        f.x = fmaf(w, dx, f.x); f.y = fmaf(w, dy, f.y); f.z = fmaf(w, dz, f.z);
    }
}

template <class P, bool POLY6, bool SOLIDS, bool LAST>
__global__ void __launch_bounds__(LGPU_TILE) k_fluid_deltap(View v, FluidParams fp, const float4* __restrict__ cur, float4* __restrict__ next) {
    extern __shared__ float4 stage[];
    __shared__ BlkDesc d;
    __shared__ uint64_t bar;
    const int i = blockIdx.x * LGPU_TILE + threadIdx.x;
    stage_begin(v, cur, d, &bar, stage);
    const int word = i < v.n ? v.nbr_cnt[i] : -1;
    if (word == -1 || (word & (LGPU_CNT_GHOST | LGPU_CNT_WALK))) stage_wait(&bar);  // (the table path waits after its row loads)
    if (word == -1) {
    } else if (word & LGPU_CNT_GHOST) {  // a neighbouring slab's particle: its owner sends the new value
        if (LAST) v.flags_in[i] = LGPU_FLAG_DEAD;
    } else {
    const float4 ci = cur[i];
    const Geom& g = v.g;
    const F3 xi = f3(ci);
    const float li = ci.w;
    const bool literal = fp.literal_lambda_index != 0;
    F3 f = f3(0.0f, 0.0f, 0.0f);
    if (!(word & LGPU_CNT_WALK)) {
        replay_neighbors<SOLIDS, false>(v, d, &bar, stage, cur, i, word & LGPU_CNT_MASK, [&](float4 pj, uint32_t code, int t) {
            // :97 — the reference indexes lambdas with the LOOP COUNTER
            float lj;
            if (literal) lj = t < LGPU_LAMBDA_HEAD ? v.lambda_head[t] : 0.0f;
            else lj = (SOLIDS && (code & LGPU_SOLID_CODE)) ? 0.0f : pj.w;
            deltap_pair<P, POLY6>(g, fp, xi, f3(pj), li, lj, f);
        });
    } else {
        int t = 0;
        walk<false>(v, i, f3(v.x0[i]), [&](int j, int) {
            float lj;
            if (literal) lj = t < LGPU_LAMBDA_HEAD ? v.lambda_head[t] : 0.0f;
            else lj = j >= 0 ? v.lambda[j] : 0.0f;
            t++;
            deltap_pair<P, POLY6>(g, fp, xi, j >= 0 ? f3(cur[j]) : f3(v.solid_pos[~j]), li, lj, f);
        });
    }
    F3 p;
    if (P::exact || POLY6) {
        f = vdiv<P>(f, fp.rest_density);                                           // :100
        p = vadd<P>(xi, f);                                                        // :103
    } else {
        p = f3(xi.x + f.x * fp.inv_rho0, xi.y + f.y * fp.inv_rho0, xi.z + f.z * fp.inv_rho0);
    }
    const float r = g.radius;
    p.x = resolve_collision(p.x, r, __fsub_rn((float)g.idomX, r));                 // :106-108
    p.y = resolve_collision(p.y, r, __fsub_rn((float)g.idomY, r));
    p.z = resolve_collision(p.z, r, __fsub_rn((float)g.idomZ, r));
    next[i] = f4(p);
    if (LAST) {
        // :110-111 (always Exact: v and x feed the next step's keys)
        // written to the step-boundary storage (the pre-reorder buffers, free since k_reorder)
        F3 xo = f3(v.pos[i]);
        v.vel_in[i] = f4(vdiv<Exact>(vsub<Exact>(p, xo), fp.dt));
        v.pos_in[i] = f4(p);
        v.flags_in[i] = v.flags[i];
        v.orig_in[i] = v.orig[i];
    }
    }
}


// ------------------------------------------------------------------------------------------
// Fast-policy kernels of the default configuration (cubic spline, lambdas[neighbour], s_corr_n = 4,
// table width 32).  Same algebra as above, arranged for the issue-bound inner loop:
//   * every list neighbour starts the substep at q <= 0.5 (the list predicate r <= h IS q <= 0.5,
//     SURVEY F3), so the loop evaluates only the inner branch of the spline, branch-free, in terms
//     of r2 and len (no q): ~20 instructions per neighbour instead of ~40;
//   * the few neighbours that have drifted beyond q = 0.5 are flagged in a bit mask and corrected
//     after the loop (true value minus what the loop added);
//   * padding entries are the particle's own slot: zero separation, every term vanishes;
//   * the thread's own loads (list length, x*, first five table groups) are issued before the
//     block-wide prologue so that they overlap the descriptor load and the bulk copies.
// Rows the table could not hold and blocks in virtual-slot mode take the generic path.
// ------------------------------------------------------------------------------------------
template <bool SOLIDS>
__device__ __forceinline__ float4 fetch_code(const View& v, const BlkDesc& d, uint32_t stage_addr, uint32_t code) {
    if (SOLIDS && (code & LGPU_SOLID_CODE)) return v.solid_pos[d.sbase[(code >> 11) & 15] + (int)(code & (LGPU_SOLID_WINDOW - 1))];
    return lds128(slot_addr(stage_addr, code));
}

// out-of-line paths of the fast kernels: virtual-slot tiles and rows the table could not hold
template <bool SOLIDS>
__device__ __noinline__ float2 lambda_slow(const View& v, const FluidParams& fp, const BlkDesc& d, uint32_t stage_addr, const float4* __restrict__ cur,
                                           int i, int word, F3 xi) {
    LambdaAcc<Fast, false> a;
    a.init();
    const Geom& g = v.g;
    if (!(word & LGPU_CNT_WALK)) {
        replay_table<SOLIDS, false>(v, d, stage_addr, cur, i, word & LGPU_CNT_MASK, [&](float4 pj, uint32_t, int) { a.pair(g, fp, xi, f3(pj)); });
    } else {
        walk<false>(v, i, f3(v.x0[i]), [&](int j, int) { a.pair(g, fp, xi, j >= 0 ? f3(cur[j]) : f3(v.solid_pos[~j])); });
    }
    const float lam = a.finish(fp);
    return make_float2(a.rho, lam);
}
template <bool SOLIDS>
__device__ __noinline__ float3 deltap_slow(const View& v, const FluidParams& fp, const BlkDesc& d, uint32_t stage_addr, const float4* __restrict__ cur,
                                           int i, int word, F3 xi, float li) {
    const Geom& g = v.g;
    F3 f = f3(0.0f, 0.0f, 0.0f);
    if (!(word & LGPU_CNT_WALK)) {
        replay_table<SOLIDS, false>(v, d, stage_addr, cur, i, word & LGPU_CNT_MASK, [&](float4 pj, uint32_t code, int) {
            const float lj = (SOLIDS && (code & LGPU_SOLID_CODE)) ? 0.0f : pj.w;
            deltap_pair<Fast, false>(g, fp, xi, f3(pj), li, lj, f);
        });
    } else {
        walk<false>(v, i, f3(v.x0[i]), [&](int j, int) {
            deltap_pair<Fast, false>(g, fp, xi, j >= 0 ? f3(cur[j]) : f3(v.solid_pos[~j]), li, j >= 0 ? v.lambda[j] : 0.0f, f);
        });
    }
    return make_float3(f.x, f.y, f.z);
}

template <bool SOLIDS>
__global__ void __launch_bounds__(LGPU_TILE, LGPU_BLOCKS_PER_SM) k_fluid_lambda_fast(const __grid_constant__ View v, const __grid_constant__ FluidParams fp, float4* __restrict__ cur) {
    extern __shared__ float4 stage[];
    __shared__ BlkDesc d;
    __shared__ uint64_t bar;
    const int i = blockIdx.x * LGPU_TILE + threadIdx.x;
    const int ic = i < v.n ? i : 0;
    pdl_trigger();
    const int word = i < v.n ? v.nbr_cnt[i] : LGPU_CNT_GHOST;
    TableRow<8> row;
    load_row_early<8, 5>(row, v, ic);
    pdl_wait();  // everything above is independent of the previous pass; x* (and lambda in its w lane) is not
    const float4 ci = cur[ic];
    stage_begin(v, cur, d, &bar, stage);
    const int cnt = word & LGPU_CNT_MASK;
    const bool table = !(word & (LGPU_CNT_GHOST | LGPU_CNT_WALK));
    if (table) load_row_rest<8, 5>(row, v, i, cnt);
    stage_wait(&bar);
    if (!(word & LGPU_CNT_GHOST)) {
    const F3 xi = f3(ci);
    const uint32_t stage_addr = smem_u32(stage);
    float rho, lam;
    if (d.mode == 0 && table) {
        float acc = 0.0f, sum = 0.0f, gx = 0.0f, gy = 0.0f, gz = 0.0f;
        uint32_t far = 0;
        replay_row<SOLIDS, true, 8, true>(v, d, row, stage_addr, cur, cnt, [&](float4 pj, uint32_t, int k) {
            const float dx = xi.x - pj.x, dy = xi.y - pj.y, dz = xi.z - pj.z;
            const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            const float len = sqrt_approx(r2);
            acc = fmaf(r2, fmaf(len, fp.fA, -fp.fB), acc);       // sum of W/cubic_k - 1
            const float gs = fmaf(len, fp.fgA, fp.gB);
            sum = fmaf(gs * gs, r2, sum);
            gx = fmaf(-gs, dx, gx); gy = fmaf(-gs, dy, gy); gz = fmaf(-gs, dz, gz);
            if (r2 > fp.thr2) far |= 1u << k;
        });
        while (far) {  // neighbours beyond q = 0.5: replace the inner-branch terms by the true ones
            const int k = __ffs(far) - 1;
            far &= far - 1;
            const float4 pj = fetch_code<SOLIDS>(v, d, stage_addr, row_code_reg(row, k));
            const float dx = xi.x - pj.x, dy = xi.y - pj.y, dz = xi.z - pj.z;
            const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            const float len = sqrt_approx(r2);
            const float wf = fmaf(r2, fmaf(len, fp.fA, -fp.fB), 1.0f);
            const float gf = fmaf(len, fp.fgA, fp.gB);
            float wt, gt;
            cubic_pair_inner(fp, r2, fp.gA, fp.gB, -fp.neg_mr * fp.l_kfh, wt, gt);
            acc += wt - wf;
            sum = fmaf(gt * gt - gf * gf, r2, sum);
            const float dg = gt - gf;
            gx = fmaf(-dg, dx, gx); gy = fmaf(-dg, dy, gy); gz = fmaf(-dg, dz, gz);
        }
        rho = fmaf(acc + (float)cnt, fp.mk, fp.mass * fp.W_zero);
        const float Ci = rho * fp.inv_rho0 - 1.0f;
        sum += gx * gx + gy * gy + gz * gz;
        lam = sum > 0.0f ? __fdividef(-Ci, sum + fp.eps) : 0.0f;
    } else {
        const float2 r = lambda_slow<SOLIDS>(v, fp, d, stage_addr, cur, i, word, xi);
        rho = r.x; lam = r.y;
    }
    v.density[i] = rho;
    v.lambda[i] = lam;
    reinterpret_cast<float*>(cur + i)[3] = lam;
    }
}

template <bool SOLIDS, bool LAST>
__global__ void __launch_bounds__(LGPU_TILE, LGPU_BLOCKS_PER_SM) k_fluid_deltap_fast(const __grid_constant__ View v, const __grid_constant__ FluidParams fp, const float4* __restrict__ cur, float4* __restrict__ next) {
    extern __shared__ float4 stage[];
    __shared__ BlkDesc d;
    __shared__ uint64_t bar;
    const int i = blockIdx.x * LGPU_TILE + threadIdx.x;
    const int ic = i < v.n ? i : 0;
    pdl_trigger();
    const int word = i < v.n ? v.nbr_cnt[i] : -1;
    TableRow<8> row;
    load_row_early<8, 5>(row, v, ic);
    pdl_wait();  // everything above is independent of the previous pass; x* and lambda are not
    const float4 ci = cur[ic];
    stage_begin(v, cur, d, &bar, stage);
    const int cnt = word & LGPU_CNT_MASK;
    const bool table = word != -1 && !(word & (LGPU_CNT_GHOST | LGPU_CNT_WALK));
    if (table) load_row_rest<8, 5>(row, v, i, cnt);
    stage_wait(&bar);
    if (word == -1) {
    } else if (word & LGPU_CNT_GHOST) {  // a neighbouring slab's particle: its owner sends the new value
        if (LAST) v.flags_in[i] = LGPU_FLAG_DEAD;
    } else {
    const Geom& g = v.g;
    const F3 xi = f3(ci);
    const float li = ci.w;
    const uint32_t stage_addr = smem_u32(stage);
    float fx = 0.0f, fy = 0.0f, fz = 0.0f;
    if (d.mode == 0 && table) {
        uint32_t far = 0;
        replay_row<SOLIDS, true, 8, true>(v, d, row, stage_addr, cur, cnt, [&](float4 pj, uint32_t code, int k) {
            const float dx = xi.x - pj.x, dy = xi.y - pj.y, dz = xi.z - pj.z;
            const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            const float len = sqrt_approx(r2);
            const float x = fmaf(r2, fmaf(len, fp.xA, -fp.xB), fp.kx);   // W / W(s_corr_dq)
            const float x2 = x * x;
            const float lj = (SOLIDS && (code & LGPU_SOLID_CODE)) ? 0.0f : pj.w;
            const float w = fmaf(-fp.s_corr_k, x2 * x2, li + lj) * fmaf(len, fp.fcA, fp.cB);
            fx = fmaf(w, dx, fx); fy = fmaf(w, dy, fy); fz = fmaf(w, dz, fz);
            if (r2 > fp.thr2) far |= 1u << k;
        });
        while (far) {
            const int k = __ffs(far) - 1;
            far &= far - 1;
            const uint32_t code = row_code_reg(row, k);
            const float4 pj = fetch_code<SOLIDS>(v, d, stage_addr, code);
            const float dx = xi.x - pj.x, dy = xi.y - pj.y, dz = xi.z - pj.z;
            const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            const float len = sqrt_approx(r2);
            const float lj = (SOLIDS && (code & LGPU_SOLID_CODE)) ? 0.0f : pj.w;
            const float xf = fmaf(r2, fmaf(len, fp.xA, -fp.xB), fp.kx);
            const float xf2 = xf * xf;
            const float wf = fmaf(-fp.s_corr_k, xf2 * xf2, li + lj) * fmaf(len, fp.fcA, fp.cB);
            float wp, cf;
            cubic_pair_inner(fp, r2, fp.cA, fp.cB, -fp.l_kfh, wp, cf);
            const float xt = wp * fp.kx;
            const float xt2 = xt * xt;
            const float dw = fmaf(-fp.s_corr_k, xt2 * xt2, li + lj) * cf - wf;
            fx = fmaf(dw, dx, fx); fy = fmaf(dw, dy, fy); fz = fmaf(dw, dz, fz);
        }
    } else {
        const float3 f = deltap_slow<SOLIDS>(v, fp, d, stage_addr, cur, i, word, xi, li);
        fx = f.x; fy = f.y; fz = f.z;
    }
    F3 p = f3(fmaf(fx, fp.inv_rho0, xi.x), fmaf(fy, fp.inv_rho0, xi.y), fmaf(fz, fp.inv_rho0, xi.z));
    const float r = g.radius;
    p.x = resolve_collision(p.x, r, __fsub_rn((float)g.idomX, r));                 // :106-108
    p.y = resolve_collision(p.y, r, __fsub_rn((float)g.idomY, r));
    p.z = resolve_collision(p.z, r, __fsub_rn((float)g.idomZ, r));
    next[i] = f4(p);
    if (LAST) {  // :110-111, always Exact (v and x feed the next step's keys); written to the step-boundary storage
        F3 xo = f3(v.pos[i]);
        v.vel_in[i] = f4(vdiv<Exact>(vsub<Exact>(p, xo), fp.dt));
        v.pos_in[i] = f4(p);
        v.flags_in[i] = v.flags[i];
        v.orig_in[i] = v.orig[i];
    }
    }
}

template <bool SOLIDS>
static int run_fluid_fast(lgpu_ctx* c, const View& v, const FluidParams& fp, int iterations) {
    const int blocks = (c->n + LGPU_TILE - 1) / LGPU_TILE > 0 ? (c->n + LGPU_TILE - 1) / LGPU_TILE : 1;
    const size_t smem = sizeof(float4) * LGPU_STAGE_SLOTS;
    float4* cur = c->x0;
    float4* bufs[2] = {c->pa, c->pb};
    const bool slab = lgpu_slab_active(c);
    // programmatic dependent launch between the passes and, in slab mode, through the refresh kernels (not with
    // per-launch event marks in between); LGPU_PDL=0 turns it off
    static const bool pdl_env = !(getenv("LGPU_PDL") && atoi(getenv("LGPU_PDL")) == 0);
    const bool pdl = pdl_env && !c->phase_timing && !c->use_graph && lgpu_slab_pdl_ok(c);
    for (int it = 0; it < iterations; it++) {
        float4* next = bufs[it & 1];
        lgpu_mark(c, 6);
        CUDA_TRY(launch_pdl(k_fluid_lambda_fast<SOLIDS>, blocks, LGPU_TILE, smem, c->stream, pdl && it > 0, v, fp, cur));
        if (slab) { lgpu_mark(c, 8); int st = lgpu_slab_refresh(c, cur, true); if (st) return st; }
        lgpu_mark(c, 7);
        if (it == iterations - 1) CUDA_TRY(launch_pdl(k_fluid_deltap_fast<SOLIDS, true>, blocks, LGPU_TILE, smem, c->stream, pdl, v, fp, (const float4*)cur, next));
        else CUDA_TRY(launch_pdl(k_fluid_deltap_fast<SOLIDS, false>, blocks, LGPU_TILE, smem, c->stream, pdl, v, fp, (const float4*)cur, next));
        c->launches += 2;
        if (slab && it < iterations - 1) { lgpu_mark(c, 8); int st = lgpu_slab_refresh(c, next, false); if (st) return st; }
        cur = next;
    }
    c->pstar_final = cur;
    CUDA_TRY(cudaGetLastError());
    return LGPU_OK;
}

template <class P, bool POLY6, bool SOLIDS>
static int run_fluid(lgpu_ctx* c, const View& v, const FluidParams& fp, int iterations) {
    const int blocks = (c->n + LGPU_TILE - 1) / LGPU_TILE > 0 ? (c->n + LGPU_TILE - 1) / LGPU_TILE : 1;
    const size_t smem = sizeof(float4) * LGPU_STAGE_SLOTS;
    float4* cur = c->x0;
    float4* bufs[2] = {c->pa, c->pb};
    const bool slab = lgpu_slab_active(c);
    for (int it = 0; it < iterations; it++) {
        float4* next = bufs[it & 1];
        lgpu_mark(c, 6);
        // slab mode: after each pass a small kernel copies the boundary particles' lambda (.w of cur) / corrected x*
        // (next) into the neighbours' ghost slots and waits for the neighbours' stores of the same pass
        k_fluid_lambda<P, POLY6, SOLIDS><<<blocks, LGPU_TILE, smem, c->stream>>>(v, fp, cur);
        if (slab && !fp.literal_lambda_index) { lgpu_mark(c, 8); int st = lgpu_slab_refresh(c, cur, true); if (st) return st; }
        lgpu_mark(c, 7);
        if (it == iterations - 1) k_fluid_deltap<P, POLY6, SOLIDS, true><<<blocks, LGPU_TILE, smem, c->stream>>>(v, fp, cur, next);
        else k_fluid_deltap<P, POLY6, SOLIDS, false><<<blocks, LGPU_TILE, smem, c->stream>>>(v, fp, cur, next);
        c->launches += 2;
        if (slab && it < iterations - 1) { lgpu_mark(c, 8); int st = lgpu_slab_refresh(c, next, false); if (st) return st; }
        cur = next;
    }
    c->pstar_final = cur;
    CUDA_TRY(cudaGetLastError());
    return LGPU_OK;
}

// host evaluation of the constants, in the reference's fp32 operation order
static float host_cubic_W(const Geom& g, float r) {
    float q = (r * g.kernel_factor) / g.h;
    float result = 0.0f;
    if (q <= 1.0) {
        if (q <= 0.5) {
            float q2 = q * q;
            float q3 = q2 * q;
            result = g.cubic_k * (6.0f * q3 - 6.0f * q2 + 1.0f);  // built with -ffp-contract=off
        } else {
            result = g.cubic_k * (2.0f * powf(1.0f - q, 3.0f));
        }
    }
    return result;
}
static float host_poly6_W(const Geom& g, float r) {
    float result = 0.0f;
    float hf = g.h * g.kernel_factor;
    if (r <= g.h) {
        double a = 315.0f / ((double)(64.0f * 3.14f) * pow((double)hf, 9.0));
        double kr = (double)(g.kernel_factor * r);
        double d = (double)hf * (double)hf - kr * kr;
        result = (float)(a * (d * d * d));
    }
    return result;
}

FluidParams lgpu_make_fluid_params(const Geom& g, const lgpu_step_params& p) {
    FluidParams fp;
    fp.dt = fminf(fmaxf(p.dt, 0.001f), 0.01f);  // src/Simulate.cpp:31
    fp.rest_density = p.rest_density; fp.mass = p.mass; fp.eps = p.relaxation_epsilon;
    fp.s_corr_k = p.s_corr_k; fp.s_corr_n = p.s_corr_n;
    const bool poly6 = p.sph_kernel == 1;
    fp.W_dq = poly6 ? host_poly6_W(g, p.s_corr_dq) : host_cubic_W(g, p.s_corr_dq);
    fp.W_zero = poly6 ? host_poly6_W(g, 0.0f) : host_cubic_W(g, 0.0f);
    fp.neg_mr = -(p.mass / p.rest_density);
    fp.c_q = g.kernel_factor / g.h;
    fp.l_h2 = g.cubic_l / (g.h * g.h);
    fp.l_kfh = g.cubic_l / (g.kernel_factor * g.h);
    fp.inv_W_dq = 1.0f / fp.W_dq;
    fp.inv_rho0 = 1.0f / p.rest_density;
    fp.inv_dt = 1.0f / fp.dt;
    fp.cA = 3.0f * fp.l_h2; fp.cB = -2.0f * fp.l_h2;
    fp.gA = fp.neg_mr * fp.cA; fp.gB = fp.neg_mr * fp.cB;
    fp.kx = g.cubic_k / fp.W_dq;
    fp.mk = p.mass * g.cubic_k;
    fp.fA = 6.0f * fp.c_q * fp.c_q * fp.c_q; fp.fB = 6.0f * fp.c_q * fp.c_q;
    fp.thr2 = (0.5f / fp.c_q) * (0.5f / fp.c_q);
    fp.fgA = fp.gA * fp.c_q; fp.fcA = fp.cA * fp.c_q;
    fp.xA = fp.kx * fp.fA; fp.xB = fp.kx * fp.fB;
    fp.literal_lambda_index = p.literal_lambda_index;
    return fp;
}

int lgpu_launch_fluid_solver(lgpu_ctx* c, const lgpu_step_params& p) {
    if (c->n == 0 && !c->slab) return LGPU_OK;
    View v = lgpu_make_view(c);
    FluidParams fp = lgpu_make_fluid_params(c->g, p);
    const int K = p.iterations < 1 ? 1 : p.iterations;
    const bool solids = c->n_solid > 0;
    if (p.sph_kernel == 1) return solids ? run_fluid<Exact, true, true>(c, v, fp, K) : run_fluid<Exact, true, false>(c, v, fp, K);
    if (p.exact_math) return solids ? run_fluid<Exact, false, true>(c, v, fp, K) : run_fluid<Exact, false, false>(c, v, fp, K);
    if (!p.literal_lambda_index && p.s_corr_n == 4.0f && c->M == 32 && !c->generic_kernels)
        return solids ? run_fluid_fast<true>(c, v, fp, K) : run_fluid_fast<false>(c, v, fp, K);
    return solids ? run_fluid<Fast, false, true>(c, v, fp, K) : run_fluid<Fast, false, false>(c, v, fp, K);
}

// ---- function tables for the kernel parity tests ----
template <class P>
__global__ void k_eval_kernel(Geom g, FluidParams fp, int which, const float* in, int n, float* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (which == 0) out[i] = cubic_W<P>(g, in[i]);
    else if (which == 2) out[i] = poly6_W(g, in[i]);
    else if (which == 4) {
        float x = P::div(cubic_W<P>(g, in[i]), fp.W_dq);
        out[i] = P::mul(-fp.s_corr_k, powf_like_libm(x, fp.s_corr_n));
    } else {
        F3 d = f3(in[3 * i], in[3 * i + 1], in[3 * i + 2]);
        F3 r;
        if (which == 1) {
            if (P::exact) r = cubic_gradW<P>(g, d);
            else {
                float r2 = d.x * d.x + d.y * d.y + d.z * d.z, Wv, coef;
                cubic_pair_fast(g, fp, r2, Wv, coef);
                r = f3(coef * d.x, coef * d.y, coef * d.z);
            }
        } else r = spiky_gradW(g, d);
        out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
    }
}

int lgpu_eval_kernel(lgpu_ctx* c, const lgpu_step_params* p, int which, const float* in, int n, float* out) {
    if (!c || !p || !in || !out || n < 0 || which < 0 || which > 4) return LGPU_ERR_ARG;
    if (n == 0) return LGPU_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    const int width = (which == 1 || which == 3) ? 3 : 1;
    const size_t words = ((size_t)n * width + 3) & ~(size_t)3;
    { int st = lgpu_scratch_reserve(c, sizeof(float) * 2 * words); if (st) return st; }
    float* d_in = (float*)c->scratch;
    float* d_out = d_in + words;
    CUDA_TRY(cudaMemcpyAsync(d_in, in, sizeof(float) * n * width, cudaMemcpyHostToDevice, c->stream));
    lgpu_step_params q = *p;
    q.sph_kernel = 0;
    FluidParams fp = lgpu_make_fluid_params(c->g, q);
    if (p->exact_math) k_eval_kernel<Exact><<<lgpu_blocks(n), LGPU_BLOCK, 0, c->stream>>>(c->g, fp, which, d_in, n, d_out);
    else k_eval_kernel<Fast><<<lgpu_blocks(n), LGPU_BLOCK, 0, c->stream>>>(c->g, fp, which, d_in, n, d_out);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, d_out, sizeof(float) * n * width, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return LGPU_OK;
}


#define LGPU_PRELOAD(f) do { cudaFuncAttributes a; CUDA_TRY(cudaFuncGetAttributes(&a, f)); } while (0)
template <class P, bool POLY6, bool SOLIDS> static int preload_fluid_variant() {
    const int smem = (int)(sizeof(float4) * LGPU_STAGE_SLOTS);
    CUDA_TRY(cudaFuncSetAttribute(k_fluid_lambda<P, POLY6, SOLIDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_TRY(cudaFuncSetAttribute(k_fluid_deltap<P, POLY6, SOLIDS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_TRY(cudaFuncSetAttribute(k_fluid_deltap<P, POLY6, SOLIDS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    LGPU_PRELOAD((k_fluid_lambda<P, POLY6, SOLIDS>));
    LGPU_PRELOAD((k_fluid_deltap<P, POLY6, SOLIDS, true>));
    LGPU_PRELOAD((k_fluid_deltap<P, POLY6, SOLIDS, false>));
    return LGPU_OK;
}
template <bool SOLIDS> static int preload_fluid_fast() {
    const int smem = (int)(sizeof(float4) * LGPU_STAGE_SLOTS);
    CUDA_TRY(cudaFuncSetAttribute(k_fluid_lambda_fast<SOLIDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_TRY(cudaFuncSetAttribute(k_fluid_deltap_fast<SOLIDS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_TRY(cudaFuncSetAttribute(k_fluid_deltap_fast<SOLIDS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    return LGPU_OK;
}
// per device (cudaFuncSetAttribute applies to the current device only): called by lgpu_create
int lgpu_preload_fluid() {
    int st = 0;
    st |= preload_fluid_fast<true>(); st |= preload_fluid_fast<false>();
    LGPU_PRELOAD(k_fluid_lambda_fast<true>); LGPU_PRELOAD(k_fluid_lambda_fast<false>);
    LGPU_PRELOAD((k_fluid_deltap_fast<true, true>)); LGPU_PRELOAD((k_fluid_deltap_fast<true, false>));
    LGPU_PRELOAD((k_fluid_deltap_fast<false, true>)); LGPU_PRELOAD((k_fluid_deltap_fast<false, false>));
    st |= preload_fluid_variant<Exact, true, true>(); st |= preload_fluid_variant<Exact, true, false>();
    st |= preload_fluid_variant<Exact, false, true>(); st |= preload_fluid_variant<Exact, false, false>();
    st |= preload_fluid_variant<Fast, false, true>(); st |= preload_fluid_variant<Fast, false, false>();
    return st ? LGPU_ERR_CUDA : LGPU_OK;
}
