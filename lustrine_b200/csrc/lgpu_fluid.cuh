// lgpu_fluid.cuh — device code of the position-based-fluids solver shared by the table build (which
// runs the FIRST density + lambda pass on the neighbours it has just found, lgpu_neighbors.cu) and the
// solver passes (lgpu_fluid.cu).  Reference: Lustrine::simulate_fluid, src/Simulate.cpp:58-113;
// cubic_kernel / cubic_kernel_grad / poly6_kernel / spiky_kernel, src/Kernels.cpp:6-67; s_coor, src/Simulate.cpp:7-9.
#pragma once
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

// Bit mask of the row entries that have drifted past q = 0.5, one bit per entry shifted in from the right: the bit is
// the sign of thr2 - r2 (set <=> r2 > thr2; the rounded difference has the exact sign and is +0 on equality), moved in
// by a funnel shift — FADD + SHF per neighbour instead of FSETP + SHF + SEL + LOP3 for `far |= (r2 > thr2) << k`.
// After a whole row of n entries (padded to groups of four), entry k sits at bit n - 1 - k.
__device__ __forceinline__ uint32_t far_push(uint32_t far, float thr2, float r2) {
    return __funnelshift_l(__float_as_uint(thr2 - r2), far, 1);
}

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


// Which density + lambda arithmetic a kernel instantiates.
//   LM_FAST      default configuration (cubic spline, fast arithmetic, lambdas[neighbour], s_corr_n = 4, table width 32):
//                branch-free inner-branch evaluation, drifted neighbours corrected after the loop
//   LM_EXACT     reference operation order, every fp32 operation separately rounded (parity mode)
//   LM_POLY6     poly6 / spiky (always exact: the reference evaluates them in double)
//   LM_GENERIC   fast arithmetic with the branches of the spline (any table width, literal lambda index)
enum { LM_NONE = 0, LM_FAST = 1, LM_EXACT = 2, LM_POLY6 = 3, LM_GENERIC = 4 };
template <int LM> struct LambdaPolicy { typedef Fast P; static constexpr bool poly6 = false; };
template <> struct LambdaPolicy<LM_EXACT> { typedef Exact P; static constexpr bool poly6 = false; };
template <> struct LambdaPolicy<LM_POLY6> { typedef Exact P; static constexpr bool poly6 = true; };

// rows the table could not hold (and every row of a brick whose neighbourhood exceeds the stage): stencil re-walk
// over the global storage with the frozen build-time predicate
template <class P, bool POLY6>
__device__ __noinline__ float2 lambda_walk(const View& v, const FluidParams& fp, const float4* cur, int i, F3 xi) {
    LambdaAcc<P, POLY6> a;
    a.init();
    const Geom& g = v.g;
    walk<false>(v, i, f3(v.x0[i]), [&](int j, int) { a.pair(g, fp, xi, j >= 0 ? f3(cur[j]) : f3(v.solid_pos[~j])); });
    const float lam = a.finish(fp);
    return make_float2(a.rho, lam);
}

// entries 32 .. cnt-1 of a list longer than the table width (spill chunk): the accumulator goes in and comes back by value
template <class P, bool POLY6>
__device__ __noinline__ LambdaAcc<P, POLY6> lambda_spill(const View& v, const FluidParams& fp, const Chunk& ck, int cnt, F3 xi, LambdaAcc<P, POLY6> acc) {
    const Geom& g = v.g;
    replay_spill<false>(v, ck.i, ck.stage_addr, cnt, [&](float4 pj, uint32_t, int) { acc.pair(g, fp, xi, f3(pj)); });
    return acc;
}

// Density + lambda of ONE particle (src/Simulate.cpp:58-88) from its table row and the staged neighbourhood.
// word = list length | LGPU_CNT_*; xi = the particle's own x*.  Writes rho_i, lambda_i, and lambda_i into the w lane of the
// particle's own x* so that the delta-p pass gets (x*_j, lambda_j) in one LDS.128.
template <int LM>
__device__ __forceinline__ void fluid_lambda_particle(const View& v, const FluidParams& fp, const Chunk& ck, float4* cur, int word, F3 xi) {
    typedef typename LambdaPolicy<LM>::P P;
    constexpr bool POLY6 = LambdaPolicy<LM>::poly6;
    const uint32_t stage_addr = ck.stage_addr;
    const int i = ck.i;
    const int cnt = word & LGPU_CNT_MASK;
    const bool table = !(word & LGPU_CNT_WALK) && ck.d->mode == 0;
    float rho, lam;
    if (LM == LM_FAST) {
        if (table) {
            float acc = 0.0f, sum = 0.0f, gx = 0.0f, gy = 0.0f, gz = 0.0f;
            uint32_t far = 0;
            const unsigned long long xi_xy = pack2(xi.x, xi.y);
            replay_row<true>(ck, cnt, [&](float4 pj, uint32_t, int k) {
                const float2 dxy = unpack2(sub2_rn(xi_xy, pack2(pj.x, pj.y)));  // (one FADD2 for the x and y lanes)
                const float dx = dxy.x, dy = dxy.y, dz = xi.z - pj.z;
                const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                const float len = sqrt_approx(r2);
                acc = fmaf(r2, fmaf(len, fp.fA, -fp.fB), acc);       // sum of W/cubic_k - 1
                const float gs = fmaf(len, fp.fgA, fp.gB);
                sum = fmaf(gs * gs, r2, sum);
                gx = fmaf(-gs, dx, gx); gy = fmaf(-gs, dy, gy); gz = fmaf(-gs, dz, gz);
                far = far_push(far, fp.thr2, r2);
            });
            const int k_top = 4 * ((min(cnt, 4 * LGPU_MG) + 3) >> 2) - 1;  // entry k sits at bit k_top - k of `far`
            while (far) {  // neighbours beyond q = 0.5: replace the inner-branch terms by the true ones
                const int b = 31 - __clz(far);
                far ^= 1u << b;
                const int k = k_top - b;
                const float4 pj = lds128(code_addr(stage_addr, row_code(ck, k)));
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
            if (cnt > 4 * LGPU_MG) {  // the tail of a long list, with the branches of the spline (LambdaAcc accumulates W/cubic_k)
                LambdaAcc<Fast, false> t;
                t.init();
                t = lambda_spill<Fast, false>(v, fp, ck, cnt, xi, t);
                acc += t.rho - (float)(cnt - 4 * LGPU_MG);
                sum += t.sum; gx += t.gi.x; gy += t.gi.y; gz += t.gi.z;
            }
            rho = fmaf(acc + (float)cnt, fp.mk, fp.mass * fp.W_zero);
            const float Ci = rho * fp.inv_rho0 - 1.0f;
            sum += gx * gx + gy * gy + gz * gz;
            lam = sum > 0.0f ? __fdividef(-Ci, sum + fp.eps) : 0.0f;
        } else {
            const float2 r = lambda_walk<Fast, false>(v, fp, cur, i, xi);
            rho = r.x; lam = r.y;
        }
    } else {
        if (table) {
            LambdaAcc<P, POLY6> acc;
            acc.init();
            const Geom& g = v.g;
            replay_row<false>(ck, cnt, [&](float4 pj, uint32_t, int) { acc.pair(g, fp, xi, f3(pj)); });
            if (cnt > 4 * LGPU_MG) acc = lambda_spill<P, POLY6>(v, fp, ck, cnt, xi, acc);
            lam = acc.finish(fp);
            rho = acc.rho;
        } else {
            const float2 r = lambda_walk<P, POLY6>(v, fp, cur, i, xi);
            rho = r.x; lam = r.y;
        }
    }
    v.density[i] = rho;
    v.lambda[i] = lam;
    reinterpret_cast<float*>(cur + i)[3] = lam;
    if (LM != LM_FAST) {
        const int o = v.orig[i];
        if (o < LGPU_LAMBDA_HEAD) v.lambda_head[o] = lam;  // lambdas[] in reference slot order, for F4
    }
}

FluidParams lgpu_make_fluid_params(const Geom& g, const lgpu_step_params& p);
// which density + lambda arithmetic the step parameters select (LM_*), for the context's table width / test hooks
int lgpu_fluid_lambda_mode(const lgpu_ctx* c, const lgpu_step_params& p);
