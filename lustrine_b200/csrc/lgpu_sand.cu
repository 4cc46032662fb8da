// lgpu_sand.cu — position-based-dynamics contact + Coulomb friction iterations for sand.
// Replaces the solver loop of Lustrine::simulate_sand, src/Simulate.cpp:226-311, and the
// velocity/position commit :316-319 (simulate_sand_credits: :401-502, different mu and the
// "no gravity" bit).
//
// One fused kernel per Jacobi iteration: contact push-out + friction + box clamp (+ commit on
// the last iteration).  Reads the current x* and the OLD positions of the neighbours, writes the
// next x* to the other ping-pong buffer (the reference's positions_tmp copy, :310).
#include "lgpu_neighbors.cuh"
#include <math.h>
#include <stdlib.h>

struct SandParams {
    float dt, mass, diameter;
    float collision_coeff, friction_coeff, mu_s, mu_k;
    float d2_contact_max;  // largest fp32 d2 with sqrt_rn(d2) <= diameter: `len > diameter` <=> d2 > this
    float cc_half;         // collision_coeff * m / (m + m), evaluated in the reference's order
    int credits;
};

__device__ __forceinline__ float rsqrt_approx(float x) { return rsqrtf(x); }

// One neighbour of the contact loop, src/Simulate.cpp:231-286.  pj = current x* of the neighbour,
// old_j() = its position at the start of the step (only evaluated for sand neighbours in contact).
template <class P, class OldJ>
__device__ __forceinline__ void sand_pair(const SandParams& sp, F3 pi, F3 xi_old, F3 pj, bool is_sand, OldJ&& old_j, F3& deltap, bool& touched) {
    // :235-240 — contact predicate, always Exact
    F3 ij = vsub<Exact>(pi, pj);
    float d2 = vdot<Exact>(ij, ij);
    if (d2 == 0.0f) {  // glm::length(ij) == 0.0f
        ij = f3(0.0f, 0.00001f, 0.0f);
        d2 = vdot<Exact>(ij, ij);
    }
    if (d2 > sp.d2_contact_max) return;  // len > particleDiameter
    touched = true;
    if (P::exact) {
        float len = __fsqrt_rn(d2);
        F3 tmp, xjdelta, nrm;
        if (is_sand) {  // :242-255
            float sc = P::mul(sp.cc_half, P::sub(len, sp.diameter));
            tmp = vdiv<P>(vscale<P>(ij, sc), len);
            xjdelta = vsub<P>(vadd<P>(pj, tmp), old_j());
            nrm = vsub<P>(vsub<P>(pi, tmp), vadd<P>(pj, tmp));
        } else {        // :268-276
            float sc = P::mul(sp.collision_coeff, P::sub(len, sp.diameter));
            tmp = vdiv<P>(vscale<P>(ij, sc), len);
            xjdelta = f3(0.0f, 0.0f, 0.0f);
            nrm = vsub<P>(vsub<P>(pi, tmp), pj);
        }
        deltap = vsub<P>(deltap, tmp);
        float d = vlen<P>(tmp);
        F3 xidelta = vsub<P>(vsub<P>(pi, tmp), xi_old);
        nrm = vnormalize<P>(nrm);
        F3 rel = vsub<P>(xidelta, xjdelta);
        F3 xtan = vsub<P>(rel, vscale<P>(nrm, vdot<P>(rel, nrm)));
        float lt = P::add(vlen<P>(xtan), 1e-9f);  // avoid0, :134
        if (P::mul(d, sp.mu_s) > lt) {
            deltap = vsub<P>(deltap, vscale<P>(xtan, sp.friction_coeff));
        } else {
            float ratio = P::div(P::mul(sp.mu_k, d), lt);
            ratio = ratio < 1.0f ? ratio : 1.0f;
            deltap = vsub<P>(deltap, vscale<P>(vscale<P>(xtan, sp.friction_coeff), ratio));
        }
    } else {
        // Fast policy: same algebra with FMA contraction and approximate rsqrt.
        float rinv = rsqrt_approx(d2);
        float len = d2 * rinv;
        float sc = (is_sand ? sp.cc_half : sp.collision_coeff) * (len - sp.diameter) * rinv;
        F3 tmp = f3(ij.x * sc, ij.y * sc, ij.z * sc);
        deltap.x -= tmp.x; deltap.y -= tmp.y; deltap.z -= tmp.z;
        float d = fabsf(sc) * len;  // |tmp|
        F3 a = f3(pi.x - tmp.x, pi.y - tmp.y, pi.z - tmp.z);
        F3 rel = f3(a.x - xi_old.x, a.y - xi_old.y, a.z - xi_old.z);
        F3 nrm;
        if (is_sand) {
            F3 b = f3(pj.x + tmp.x, pj.y + tmp.y, pj.z + tmp.z);
            F3 xo = old_j();
            rel.x -= b.x - xo.x; rel.y -= b.y - xo.y; rel.z -= b.z - xo.z;
            nrm = f3(a.x - b.x, a.y - b.y, a.z - b.z);
        } else {
            nrm = f3(a.x - pj.x, a.y - pj.y, a.z - pj.z);
        }
        float ninv = rsqrt_approx(nrm.x * nrm.x + nrm.y * nrm.y + nrm.z * nrm.z);
        nrm.x *= ninv; nrm.y *= ninv; nrm.z *= ninv;
        float dn = rel.x * nrm.x + rel.y * nrm.y + rel.z * nrm.z;
        F3 xtan = f3(rel.x - dn * nrm.x, rel.y - dn * nrm.y, rel.z - dn * nrm.z);
        float lt = sqrtf(xtan.x * xtan.x + xtan.y * xtan.y + xtan.z * xtan.z) + 1e-9f;
        float s = sp.friction_coeff;
        if (!(d * sp.mu_s > lt)) s *= fminf(__fdividef(sp.mu_k * d, lt), 1.0f);
        deltap.x -= s * xtan.x; deltap.y -= s * xtan.y; deltap.z -= s * xtan.z;
    }
}

// entries 32 .. cnt-1 of a list longer than the table width (spill chunk): returns the displacement, w != 0 if touched
// (everything by value: a Chunk passed by reference to an out-of-line function would have to live in local memory)
template <class P>
__device__ __noinline__ float4 sand_spill(const View& v, const SandParams& sp, int i, uint32_t stage_addr, uint32_t solid_base, int cnt, F3 pi, F3 xi_old,
                                          F3 deltap, bool touched) {
    replay_spill<true>(v, i, stage_addr, cnt, [&](float4 pj, uint32_t code, int) {
        const bool is_sand = code < solid_base;
        sand_pair<P>(sp, pi, xi_old, f3(pj), is_sand, [&]() { return f3(v.pos[__float_as_int(pj.w)]); }, deltap, touched);
    });
    return make_float4(deltap.x, deltap.y, deltap.z, touched ? 1.0f : 0.0f);
}

template <class P, bool LAST>
__global__ void __launch_bounds__(LGPU_BRICK_THREADS, LGPU_CTAS_PER_SM) k_sand_iteration(const __grid_constant__ View v, const __grid_constant__ SandParams sp,
                                                                          const float4* cur, float4* next, int* cursor) {
    extern __shared__ unsigned char smem_raw[];
    brick_loop<false>(v, cur, cursor, smem_raw, [&](const Chunk& ck) -> int {
        const int i = ck.i, word = ck.word;
        const int mode = ck.d->mode;
        if (word & (LGPU_CNT_GHOST | LGPU_CNT_GHOST_INNER)) {  // a neighbouring slab's particle: its owner sends the new value
            if (LAST) v.flags_in[i] = LGPU_FLAG_DEAD;
            return 0;
        }
        const Geom& g = v.g;
        const uint32_t stage_addr = ck.stage_addr;
        const float4 ci = mode == 0 ? lds128(slot_addr(stage_addr, (uint32_t)ck.slot)) : cur[i];
        const F3 pi = f3(ci);
        const F3 xi_old = f3(v.pos[i]);
        F3 deltap = f3(0.0f, 0.0f, 0.0f);
        bool touched = false;
        if (!(word & LGPU_CNT_WALK) && mode == 0) {
            // Two sweeps over the row.  Only a third of a list's entries are in contact (on a packing at rest: the 6
            // nearest of 18 neighbours within h), but a warp runs the long contact body whenever ANY of its lanes needs
            // it — i.e. for every entry.  So the first sweep only evaluates the contact predicate (the reference's,
            // src/Simulate.cpp:235-240, Exact) into a bit mask, and the second runs the contact + friction body over the
            // set bits: max-over-lanes(contacts) trips instead of the full row.  Same entries, same order, same
            // arithmetic: entries that are not in contact never touched deltap.  The old position of a sand neighbour
            // in contact is read from the sorted storage at the slot its staged x* carries in the w lane.
            const uint32_t solid_base = (uint32_t)ck.d->solid_base << 4;  // (table codes are stage slots x 16)
            uint32_t contact = 0;
            const unsigned long long pi_xy = pack2(pi.x, pi.y);
            // (`apart` collects the sign of d2_contact_max - d2 per entry, shifted in from the right by a funnel shift: set
            // <=> d2 > d2_contact_max, exactly — FADD + SHF per entry; entry k ends up at bit k_top - k)
            uint32_t apart = 0;
            replay_row<true>(ck, word & LGPU_CNT_MASK, [&](float4 pj, uint32_t, int) {
                apart = __funnelshift_l(__float_as_uint(__fsub_rn(sp.d2_contact_max, dist2_exact(pi_xy, pi.z, pj))), apart, 1);
            });
            const int n_row = 4 * ((min(word & LGPU_CNT_MASK, 4 * LGPU_MG) + 3) >> 2), k_top = n_row - 1;
            contact = ~apart & (n_row ? 0xffffffffu >> (32 - n_row) : 0u);
            while (contact) {  // ascending entry = descending bit: the reference's list order
                const int b = 31 - __clz(contact);
                contact ^= 1u << b;
                const int k = k_top - b;
                const uint32_t code = row_code(ck, k);
                const float4 pj = lds128(code_addr(stage_addr, code));
                const bool is_sand = code < solid_base;
                sand_pair<P>(sp, pi, xi_old, f3(pj), is_sand, [&]() { return f3(v.pos[__float_as_int(pj.w)]); }, deltap, touched);
            }
            if ((word & LGPU_CNT_MASK) > 4 * LGPU_MG) {
                const float4 r = sand_spill<P>(v, sp, i, stage_addr, solid_base, word & LGPU_CNT_MASK, pi, xi_old, deltap, touched);
                deltap = f3(r.x, r.y, r.z); touched = r.w != 0.0f;
            }
        } else {
            walk<true>(v, i, f3(v.x0[i]), [&](int j, int) {
                const bool is_sand = j >= 0;
                sand_pair<P>(sp, pi, xi_old, is_sand ? f3(cur[j]) : f3(v.solid_pos[~j]), is_sand, [&]() { return f3(v.pos[j]); }, deltap, touched);
            });
        }
        F3 ps = P::exact ? vadd<Exact>(pi, deltap) : f3(pi.x + deltap.x, pi.y + deltap.y, pi.z + deltap.z);  // :288
        const float r = g.radius;
        ps.x = fminf(fmaxf(ps.x, r), __fsub_rn(g.domainX, r));  // :307
        ps.y = fminf(fmaxf(ps.y, r), __fsub_rn(g.domainY, r));
        ps.z = fminf(fmaxf(ps.z, r), __fsub_rn(g.domainZ, r));
        next[i] = f4(ps, __int_as_float(i));  // w = own sorted slot (see the replay above)
        if (sp.credits && touched) {  // :463-470: the "no gravity" bit drops at the first contact
            int a = v.flags[i];
            if (a & 2) v.flags[i] = a & ~2;
        }
        if (LAST) {  // :316-319, always Exact
            // Written to the step-boundary storage (the pre-reorder buffers, free since k_reorder):
            // other threads still read the OLD sorted positions v.pos[j] in this launch.
            v.vel_in[i] = f4(vdiv<Exact>(vsub<Exact>(ps, xi_old), sp.dt));
            v.pos_in[i] = f4(ps);
            v.flags_in[i] = v.flags[i];
            v.orig_in[i] = v.orig[i];
        }
        return 0;
    });
}

// largest fp32 x with sqrt_rn(x) <= d  (so that `sqrt(d2) > d` <=> `d2 > x`, bit-exactly)
static float contact_threshold(float d) {
    float x = d * d;
    while (sqrtf(x) > d) x = nextafterf(x, 0.0f);
    while (sqrtf(nextafterf(x, INFINITY)) <= d) x = nextafterf(x, INFINITY);
    return x;
}

int lgpu_launch_sand_solver(lgpu_ctx* c, const lgpu_step_params& p) {
    if (c->n == 0 && !c->slab) return LGPU_OK;
    View v = lgpu_make_view(c);
    SandParams sp;
    sp.dt = p.dt; sp.mass = p.mass; sp.diameter = c->g.diameter;
    sp.collision_coeff = p.collision_coeff; sp.friction_coeff = p.friction_coeff;
    sp.mu_s = p.mu_s; sp.mu_k = p.mu_k;
    sp.d2_contact_max = contact_threshold(c->g.diameter);
    sp.cc_half = p.collision_coeff * p.mass / (p.mass + p.mass);  // src/Simulate.cpp:246, left to right
    sp.credits = p.credits;
    int K = p.iterations < 1 ? 1 : p.iterations;
    if (1 + K > LGPU_MAX_PASSES) K = LGPU_MAX_PASSES - 1;
    const int grid = LGPU_CTAS_PER_SM * c->num_sms;
    const size_t smem = LGPU_BRICK_SMEM;
    const float4* cur = c->x0;
    float4* bufs[2] = {c->pa, c->pb};
    const bool pdl = lgpu_pdl_enabled(c) && lgpu_slab_pdl_ok(c);  // see run_fluid
    for (int it = 0; it < K; it++) {
        float4* next = bufs[it & 1];
        const bool last = it == K - 1;
        lgpu_mark(c, 7);
        if (p.exact_math) {
            if (last) CUDA_TRY(launch_pdl(k_sand_iteration<Exact, true>, grid, LGPU_BRICK_THREADS, smem, c->stream, pdl, v, sp, cur, next, c->brick_ctl + 8 + c->pass));
            else CUDA_TRY(launch_pdl(k_sand_iteration<Exact, false>, grid, LGPU_BRICK_THREADS, smem, c->stream, pdl, v, sp, cur, next, c->brick_ctl + 8 + c->pass));
        } else {
            if (last) CUDA_TRY(launch_pdl(k_sand_iteration<Fast, true>, grid, LGPU_BRICK_THREADS, smem, c->stream, pdl, v, sp, cur, next, c->brick_ctl + 8 + c->pass));
            else CUDA_TRY(launch_pdl(k_sand_iteration<Fast, false>, grid, LGPU_BRICK_THREADS, smem, c->stream, pdl, v, sp, cur, next, c->brick_ctl + 8 + c->pass));
        }
        c->pass++; c->launches++;
        if (!last && lgpu_slab_active(c)) { lgpu_mark(c, 8); int st = lgpu_slab_refresh(c, next, false); if (st) return st; }
        cur = next;
    }
    c->pstar_final = (float4*)cur;
    CUDA_TRY(cudaGetLastError());
    return LGPU_OK;
}


#define LGPU_PRELOAD(f) do { cudaFuncAttributes a; CUDA_TRY(cudaFuncGetAttributes(&a, f)); } while (0)
template <class P> static int preload_sand_variant() {
    const int smem = (int)LGPU_BRICK_SMEM;
    CUDA_TRY(cudaFuncSetAttribute((k_sand_iteration<P, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_TRY(cudaFuncSetAttribute((k_sand_iteration<P, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    return LGPU_OK;
}
int lgpu_preload_sand() {
    int st = preload_sand_variant<Exact>() | preload_sand_variant<Fast>();
    return st ? LGPU_ERR_CUDA : LGPU_OK;
}
