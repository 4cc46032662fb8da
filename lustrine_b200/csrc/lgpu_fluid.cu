// lgpu_fluid.cu — position-based-fluids solver iterations.
// Replaces the two solver loops of Lustrine::simulate_fluid, src/Simulate.cpp:58-88 (density +
// lambda) and :90-113 (delta-p, box collision, velocity/position commit), with s_coor (:7-9),
// resolve_collision (:13-24), cubic_kernel / cubic_kernel_grad (src/Kernels.cpp:6-41) inlined.
//
// Two fused kernels per solver iteration, both persistent brick kernels (lgpu_neighbors.cuh):
//   k_fluid_lambda   reads x* of the neighbours, writes rho_i and lambda_i          (16 B/particle)
//                    (the FIRST iteration's pass runs inside the table build, lgpu_neighbors.cu)
//   k_fluid_deltap   reads x*, lambda of the neighbours, writes the corrected x*    (28 B/particle)
//                    and on the last iteration also v and x                         (+24 B/particle)
// The delta-p output is double-buffered (Jacobi); the reference's in-place loop is sequential
// Gauss-Seidel in index order and is compared through the Jacobi oracle.
#include <stdlib.h>

#include "lgpu_fluid.cuh"

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


// ---- density + lambda, iterations after the first ----
template <int LM>
__global__ void __launch_bounds__(LGPU_BRICK_THREADS, LGPU_CTAS_PER_SM) k_fluid_lambda(const __grid_constant__ View v, const __grid_constant__ FluidParams fp,
                                                                        float4* cur, int* cursor) {
    extern __shared__ unsigned char smem_raw[];
    brick_loop<false>(v, cur, cursor, smem_raw, [&](const Chunk& ck) -> int {
        if (ck.word & LGPU_CNT_GHOST) return 0;
        const float4 ci = ck.d->mode == 0 ? lds128(slot_addr(ck.stage_addr, (uint32_t)ck.slot)) : cur[ck.i];
        fluid_lambda_particle<LM>(v, fp, ck, cur, ck.word, f3(ci));
        return 0;
    });
}

// rows the table could not hold: stencil re-walk over the global storage
template <class P, bool POLY6>
__device__ __noinline__ float3 deltap_walk(const View& v, const FluidParams& fp, const float4* cur, int i, F3 xi, float li) {
    const Geom& g = v.g;
    const bool literal = fp.literal_lambda_index != 0;
    F3 f = f3(0.0f, 0.0f, 0.0f);
    int t = 0;
    walk<false>(v, i, f3(v.x0[i]), [&](int j, int) {
        float lj;
        if (literal) lj = t < LGPU_LAMBDA_HEAD ? v.lambda_head[t] : 0.0f;  // :97 — lambdas[LOOP COUNTER]
        else lj = j >= 0 ? v.lambda[j] : 0.0f;
        t++;
        deltap_pair<P, POLY6>(g, fp, xi, j >= 0 ? f3(cur[j]) : f3(v.solid_pos[~j]), li, lj, f);
    });
    return make_float3(f.x, f.y, f.z);
}

// entries 32 .. cnt-1 of a list longer than the table width (spill chunk)
template <class P, bool POLY6>
__device__ __noinline__ float3 deltap_spill(const View& v, const FluidParams& fp, const Chunk& ck, int cnt, F3 xi, float li, float3 f0) {
    const Geom& g = v.g;
    const bool literal = fp.literal_lambda_index != 0;
    F3 f = f3(f0.x, f0.y, f0.z);
    replay_spill<false>(v, ck.i, ck.stage_addr, cnt, [&](float4 pj, uint32_t, int t) {
        const float lj = literal ? (t < LGPU_LAMBDA_HEAD ? v.lambda_head[t] : 0.0f) : pj.w;
        deltap_pair<P, POLY6>(g, fp, xi, f3(pj), li, lj, f);
    });
    return make_float3(f.x, f.y, f.z);
}

// ---- delta-p + box collision (+ commit on the last iteration) ----
// The staged neighbourhood holds (x*_j, lambda_j) per slot (the lambda pass stored lambda in the w lane; the w lane
// of a solid is 0, which is the lambda the reference reads for it).
template <int LM, bool LAST>
__global__ void __launch_bounds__(LGPU_BRICK_THREADS, LGPU_CTAS_PER_SM) k_fluid_deltap(const __grid_constant__ View v, const __grid_constant__ FluidParams fp,
                                                                        const float4* cur, float4* next, int* cursor) {
    typedef typename LambdaPolicy<LM>::P P;
    constexpr bool POLY6 = LambdaPolicy<LM>::poly6;
    extern __shared__ unsigned char smem_raw[];
    brick_loop<false>(v, cur, cursor, smem_raw, [&](const Chunk& ck) -> int {
        const int i = ck.i, word = ck.word, slot = ck.slot;
        const int mode = ck.d->mode;
        if (word & (LGPU_CNT_GHOST | LGPU_CNT_GHOST_INNER)) {  // a neighbouring slab's particle: its owner sends the new position
            if (LAST) v.flags_in[i] = LGPU_FLAG_DEAD;
            return 0;
        }
        const Geom& g = v.g;
        const uint32_t stage_addr = ck.stage_addr;
        const int cnt = word & LGPU_CNT_MASK;
        const bool table = !(word & LGPU_CNT_WALK) && mode == 0;
        float4 xo4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        int fl = 0, og = 0;
        if (LAST) { xo4 = v.pos[i]; fl = v.flags[i]; og = v.orig[i]; }  // (issued before the gather: the latency hides behind it)
        float fx = 0.0f, fy = 0.0f, fz = 0.0f;
        F3 xi;
        if (LM == LM_FAST) {
            const float4 ci = mode == 0 ? lds128(slot_addr(stage_addr, (uint32_t)slot)) : cur[i];
            xi = f3(ci);
            const float li = ci.w;
            if (table) {
                uint32_t far = 0;
                const unsigned long long xi_xy = pack2(xi.x, xi.y);
                replay_row<true>(ck, cnt, [&](float4 pj, uint32_t, int k) {
                    const float2 dxy = unpack2(sub2_rn(xi_xy, pack2(pj.x, pj.y)));  // (one FADD2 for the x and y lanes)
                    const float dx = dxy.x, dy = dxy.y, dz = xi.z - pj.z;
                    const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    const float len = sqrt_approx(r2);
                    const float x = fmaf(r2, fmaf(len, fp.xA, -fp.xB), fp.kx);   // W / W(s_corr_dq)
                    const float x2 = x * x;
                    const float w = fmaf(-fp.s_corr_k, x2 * x2, li + pj.w) * fmaf(len, fp.fcA, fp.cB);
                    fx = fmaf(w, dx, fx); fy = fmaf(w, dy, fy); fz = fmaf(w, dz, fz);
                    far = far_push(far, fp.thr2, r2);
                });
                const int k_top = 4 * ((min(cnt, 4 * LGPU_MG) + 3) >> 2) - 1;  // entry k sits at bit k_top - k of `far`
                while (far) {  // neighbours beyond q = 0.5: replace the inner-branch term by the true one
                    const int b = 31 - __clz(far);
                    far ^= 1u << b;
                    const int k = k_top - b;
                    const float4 pj = lds128(code_addr(stage_addr, row_code(ck, k)));
                    const float dx = xi.x - pj.x, dy = xi.y - pj.y, dz = xi.z - pj.z;
                    const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    const float len = sqrt_approx(r2);
                    const float xf = fmaf(r2, fmaf(len, fp.xA, -fp.xB), fp.kx);
                    const float xf2 = xf * xf;
                    const float wf = fmaf(-fp.s_corr_k, xf2 * xf2, li + pj.w) * fmaf(len, fp.fcA, fp.cB);
                    float wp, cf;
                    cubic_pair_inner(fp, r2, fp.cA, fp.cB, -fp.l_kfh, wp, cf);
                    const float xt = wp * fp.kx;
                    const float xt2 = xt * xt;
                    const float dw = fmaf(-fp.s_corr_k, xt2 * xt2, li + pj.w) * cf - wf;
                    fx = fmaf(dw, dx, fx); fy = fmaf(dw, dy, fy); fz = fmaf(dw, dz, fz);
                }
                if (cnt > 4 * LGPU_MG) {
                    const float3 f = deltap_spill<Fast, false>(v, fp, ck, cnt, xi, li, make_float3(fx, fy, fz));
                    fx = f.x; fy = f.y; fz = f.z;
                }
            } else {
                const float3 f = deltap_walk<Fast, false>(v, fp, cur, i, xi, li);
                fx = f.x; fy = f.y; fz = f.z;
            }
        } else {
            const float4 ci = mode == 0 ? lds128(slot_addr(stage_addr, (uint32_t)slot)) : cur[i];
            xi = f3(ci);
            const float li = ci.w;
            if (table) {
                const bool literal = fp.literal_lambda_index != 0;
                F3 f = f3(0.0f, 0.0f, 0.0f);
                replay_row<false>(ck, cnt, [&](float4 pj, uint32_t, int t) {
                    // :97 — the reference indexes lambdas with the LOOP COUNTER
                    const float lj = literal ? (t < LGPU_LAMBDA_HEAD ? v.lambda_head[t] : 0.0f) : pj.w;
                    deltap_pair<P, POLY6>(g, fp, xi, f3(pj), li, lj, f);
                });
                fx = f.x; fy = f.y; fz = f.z;
                if (cnt > 4 * LGPU_MG) {
                    const float3 f2 = deltap_spill<P, POLY6>(v, fp, ck, cnt, xi, li, make_float3(fx, fy, fz));
                    fx = f2.x; fy = f2.y; fz = f2.z;
                }
            } else {
                const float3 f = deltap_walk<P, POLY6>(v, fp, cur, i, xi, li);
                fx = f.x; fy = f.y; fz = f.z;
            }
        }
        F3 p;
        if (P::exact || POLY6) {
            const F3 f = vdiv<P>(f3(fx, fy, fz), fp.rest_density);                       // :100
            p = vadd<P>(xi, f);                                                         // :103
        } else {
            p = f3(fmaf(fx, fp.inv_rho0, xi.x), fmaf(fy, fp.inv_rho0, xi.y), fmaf(fz, fp.inv_rho0, xi.z));
        }
        const float r = g.radius;
        p.x = resolve_collision(p.x, r, __fsub_rn((float)g.idomX, r));                 // :106-108
        p.y = resolve_collision(p.y, r, __fsub_rn((float)g.idomY, r));
        p.z = resolve_collision(p.z, r, __fsub_rn((float)g.idomZ, r));
        next[i] = f4(p);
        if (LAST) {  // :110-111, always Exact (v and x feed the next step's keys); written to the step-boundary storage
            v.vel_in[i] = f4(vdiv<Exact>(vsub<Exact>(p, f3(xo4)), fp.dt));
            v.pos_in[i] = f4(p);
            v.flags_in[i] = fl;
            v.orig_in[i] = og;
        }
        return 0;
    });
}

int lgpu_fluid_lambda_mode(const lgpu_ctx* c, const lgpu_step_params& p) {
    if (p.sph_kernel == 1) return LM_POLY6;
    if (p.exact_math) return LM_EXACT;
    if (!p.literal_lambda_index && p.s_corr_n == 4.0f && c->M == 32 && !c->generic_kernels) return LM_FAST;
    return LM_GENERIC;
}

template <int LM>
static int run_fluid(lgpu_ctx* c, const View& v, const FluidParams& fp, int iterations) {
    const int grid = LGPU_CTAS_PER_SM * c->num_sms;
    const size_t smem = LGPU_BRICK_SMEM;
    float4* cur = c->x0;
    float4* bufs[2] = {c->pa, c->pb};
    const bool slab = lgpu_slab_active(c);
    // programmatic dependent launch between the passes and, in slab mode, through the refresh kernels (not with
    // per-launch event marks in between); LGPU_PDL=0 turns it off
    const bool pdl = lgpu_pdl_enabled(c) && lgpu_slab_pdl_ok(c);
    for (int it = 0; it < iterations; it++) {
        float4* next = bufs[it & 1];
        if (it > 0) {  // (the first density + lambda pass ran inside the table build)
            lgpu_mark(c, 6);
            CUDA_TRY(launch_pdl(k_fluid_lambda<LM>, grid, LGPU_BRICK_THREADS, smem, c->stream, pdl, v, fp, cur, c->brick_ctl + 8 + c->pass));
            c->pass++; c->launches++;
        }
        // slab mode: after each pass a small kernel copies the boundary particles' lambda (.w of cur) / corrected x*
        // (next) into the neighbours' ghost slots and waits for the neighbours' stores of the same pass
        // (with a two-column ghost layer the ghosts whose lambda the owned particles read have computed it themselves)
        if (slab && !fp.literal_lambda_index && c->g.gw < 2) { lgpu_mark(c, 8); int st = lgpu_slab_refresh(c, cur, true); if (st) return st; }
        lgpu_mark(c, 7);
        if (it == iterations - 1) CUDA_TRY(launch_pdl(k_fluid_deltap<LM, true>, grid, LGPU_BRICK_THREADS, smem, c->stream, pdl, v, fp, (const float4*)cur, next, c->brick_ctl + 8 + c->pass));
        else CUDA_TRY(launch_pdl(k_fluid_deltap<LM, false>, grid, LGPU_BRICK_THREADS, smem, c->stream, pdl, v, fp, (const float4*)cur, next, c->brick_ctl + 8 + c->pass));
        c->pass++; c->launches++;
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
    int K = p.iterations < 1 ? 1 : p.iterations;
    if (1 + 2 * K > LGPU_MAX_PASSES) K = (LGPU_MAX_PASSES - 1) / 2;
    switch (lgpu_fluid_lambda_mode(c, p)) {
        case LM_POLY6: return run_fluid<LM_POLY6>(c, v, fp, K);
        case LM_EXACT: return run_fluid<LM_EXACT>(c, v, fp, K);
        case LM_FAST: return run_fluid<LM_FAST>(c, v, fp, K);
        default: return run_fluid<LM_GENERIC>(c, v, fp, K);
    }
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
template <int LM> static int preload_fluid_variant() {
    const int smem = (int)LGPU_BRICK_SMEM;
    CUDA_TRY(cudaFuncSetAttribute(k_fluid_lambda<LM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_TRY(cudaFuncSetAttribute((k_fluid_deltap<LM, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_TRY(cudaFuncSetAttribute((k_fluid_deltap<LM, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    return LGPU_OK;
}
// per device (cudaFuncSetAttribute applies to the current device only): called by lgpu_create
int lgpu_preload_fluid() {
    int st = preload_fluid_variant<LM_FAST>() | preload_fluid_variant<LM_EXACT>() | preload_fluid_variant<LM_POLY6>() | preload_fluid_variant<LM_GENERIC>();
    LGPU_PRELOAD(k_eval_kernel<Exact>); LGPU_PRELOAD(k_eval_kernel<Fast>);
    return st ? LGPU_ERR_CUDA : LGPU_OK;
}
