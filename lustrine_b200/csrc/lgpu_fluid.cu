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

template <class P, bool POLY6>
__global__ void __launch_bounds__(LGPU_BLOCK) k_fluid_lambda(View v, FluidParams fp, const float4* __restrict__ cur) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.n_owned) return;
    const Geom& g = v.g;
    F3 xi = f3(cur[i]);
    float rho = 0.0f, sum = 0.0f;
    F3 gi = f3(0.0f, 0.0f, 0.0f);
    for_each_neighbor<false>(v, i, [&](int j) {
        F3 xj = j >= 0 ? f3(cur[j]) : f3(v.solid_pos[~j]);
        if (P::exact || POLY6) {
            F3 d = vsub<P>(xi, xj);
            float len = vlen<P>(d);
            rho = P::add(rho, P::mul(fp.mass, W_of<POLY6, P>(g, len)));           // :62-64
            F3 gr = vscale<P>(gradW_of<POLY6, P>(g, d), fp.neg_mr);                // :76
            sum = P::add(sum, vdot<P>(gr, gr));                                    // :77
            gi = vsub<P>(gi, gr);                                                  // :78
        } else {
            float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
            float r2 = dx * dx + dy * dy + dz * dz;
            float Wv, coef;
            cubic_pair_fast(g, fp, r2, Wv, coef);
            rho += fp.mass * Wv;
            float gs = fp.neg_mr * coef;
            sum += gs * gs * r2;
            gi.x -= gs * dx; gi.y -= gs * dy; gi.z -= gs * dz;
        }
    });
    float lam = 0.0f;
    if (P::exact || POLY6) {
        rho = P::add(rho, P::mul(fp.mass, fp.W_zero));                             // :66
        float Ci = P::sub(P::div(rho, fp.rest_density), 1.0f);                     // :69
        sum = P::add(sum, vdot<P>(gi, gi));                                        // :81
        if (sum > 0.0f) lam = P::div(-Ci, P::add(sum, fp.eps));                    // :83-86
    } else {
        rho += fp.mass * fp.W_zero;
        float Ci = rho * fp.inv_rho0 - 1.0f;
        sum += gi.x * gi.x + gi.y * gi.y + gi.z * gi.z;
        if (sum > 0.0f) lam = __fdividef(-Ci, sum + fp.eps);
    }
    v.density[i] = rho;
    v.lambda[i] = lam;
    int o = v.orig[i];
    if (o < LGPU_LAMBDA_HEAD) v.lambda_head[o] = lam;  // lambdas[] in reference slot order, for F4
}

// resolve_collision, src/Simulate.cpp:13-24 (returns 0.01, not min)
__device__ __forceinline__ float resolve_collision(float value, float lo, float hi) {
    if (value <= lo) return 0.01f;
    if (value > hi) return __fsub_rn(hi, 0.01f);
    return value;
}

template <class P, bool POLY6, bool LAST>
__global__ void __launch_bounds__(LGPU_BLOCK) k_fluid_deltap(View v, FluidParams fp, const float4* __restrict__ cur, float4* __restrict__ next) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.n_owned) return;
    const Geom& g = v.g;
    F3 xi = f3(cur[i]);
    const float li = v.lambda[i];
    F3 f = f3(0.0f, 0.0f, 0.0f);
    int t = 0;
    for_each_neighbor<false>(v, i, [&](int j) {
        F3 xj = j >= 0 ? f3(cur[j]) : f3(v.solid_pos[~j]);
        // :97 — the reference indexes lambdas with the LOOP COUNTER
        float lj;
        if (fp.literal_lambda_index) lj = t < LGPU_LAMBDA_HEAD ? v.lambda_head[t] : 0.0f;
        else lj = j >= 0 ? v.lambda[j] : 0.0f;
        t++;
        if (P::exact || POLY6) {
            F3 d = vsub<P>(xi, xj);
            float len = vlen<P>(d);
            float x = P::div(W_of<POLY6, P>(g, len), fp.W_dq);
            float sc = P::mul(-fp.s_corr_k, powf_like_libm(x, fp.s_corr_n));     // :7-9
            float w = P::add(P::add(li, lj), sc);
            f = vadd<P>(f, vscale<P>(gradW_of<POLY6, P>(g, d), w));               // :97
        } else {
            float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
            float r2 = dx * dx + dy * dy + dz * dz;
            float Wv, coef;
            cubic_pair_fast(g, fp, r2, Wv, coef);
            float x = Wv * fp.inv_W_dq;
            float x2 = x * x;
            float pw = fp.s_corr_n == 4.0f ? x2 * x2 : __powf(x, fp.s_corr_n);
            float w = (li + lj - fp.s_corr_k * pw) * coef;
            f.x += w * dx; f.y += w * dy; f.z += w * dz;
        }
    });
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

template <class P, bool POLY6>
static int run_fluid(lgpu_ctx* c, const View& v, const FluidParams& fp, int iterations) {
    const int blocks = lgpu_blocks(c->n_owned);
    const float4* cur = c->x0;
    float4* bufs[2] = {c->pa, c->pb};
    for (int it = 0; it < iterations; it++) {
        float4* next = bufs[it & 1];
        lgpu_mark(c, 6);
        k_fluid_lambda<P, POLY6><<<blocks, LGPU_BLOCK, 0, c->stream>>>(v, fp, cur);
        lgpu_mark(c, 7);
        if (it == iterations - 1) k_fluid_deltap<P, POLY6, true><<<blocks, LGPU_BLOCK, 0, c->stream>>>(v, fp, cur, next);
        else k_fluid_deltap<P, POLY6, false><<<blocks, LGPU_BLOCK, 0, c->stream>>>(v, fp, cur, next);
        c->launches += 2;
        cur = next;
    }
    c->pstar_final = (float4*)cur;
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
    fp.literal_lambda_index = p.literal_lambda_index;
    return fp;
}

int lgpu_launch_fluid_solver(lgpu_ctx* c, const lgpu_step_params& p) {
    if (c->n_owned == 0) return LGPU_OK;
    View v = lgpu_make_view(c);
    FluidParams fp = lgpu_make_fluid_params(c->g, p);
    const int K = p.iterations < 1 ? 1 : p.iterations;
    if (p.sph_kernel == 1) return run_fluid<Exact, true>(c, v, fp, K);
    if (p.exact_math) return run_fluid<Exact, false>(c, v, fp, K);
    return run_fluid<Fast, false>(c, v, fp, K);
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
    float *d_in, *d_out;
    CUDA_TRY(cudaMalloc(&d_in, sizeof(float) * n * width));
    CUDA_TRY(cudaMalloc(&d_out, sizeof(float) * n * width));
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
    cudaFree(d_in); cudaFree(d_out);
    return LGPU_OK;
}
