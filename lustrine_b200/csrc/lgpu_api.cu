// lgpu_api.cu — context lifetime, state transfer, step drivers, stage dumps (C ABI of include/lgpu.h).
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "lgpu_fluid.cuh"

static thread_local char g_err[512] = "";
void lgpu_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char* lgpu_last_error(void) { return g_err; }

extern "C" void lgpu_default_step_params(lgpu_step_params* p) {
    memset(p, 0, sizeof(*p));
    p->dt = 0.01f;
    p->gravity[0] = 0.0f; p->gravity[1] = -10.0f; p->gravity[2] = 0.0f;  // src/Simulation.hpp:162
    p->rest_density = 24.0f; p->mass = 5.0f; p->relaxation_epsilon = 10.0f;  // :159-160,165
    p->s_corr_dq = 0.5f; p->s_corr_k = 1.0f; p->s_corr_n = 4.0f;            // :168-170
    p->iterations = 1;
    p->literal_lambda_index = 1;
    p->exact_math = 1;
    p->sph_kernel = 0;
    p->attract_radius = 1.5f; p->blow_radius = 2.0f; p->attract_coeff = 1000.0f; p->blow_coeff = 500.0f;  // :229-232
    p->collision_coeff = 0.8f; p->friction_coeff = 0.7f; p->mu_s = 0.95f; p->mu_k = 0.9f;  // src/Simulate.cpp:159-163
    p->credits = 0;
}

static const double kPi = 3.14159265358979323846;  // src/Lustrine.cpp:16

static void make_geom(const lgpu_config& cfg, Geom* g) {
    // src/Lustrine.cpp:105-107,251-266, same fp32/double operation order
    g->domainX = (float)cfg.domain[0]; g->domainY = (float)cfg.domain[1]; g->domainZ = (float)cfg.domain[2];
    g->idomX = (int)g->domainX; g->idomY = (int)g->domainY; g->idomZ = (int)g->domainZ;
    g->radius = cfg.particle_radius; g->diameter = cfg.particle_diameter;
    g->h = cfg.kernel_radius_scale * cfg.particle_radius;
    g->h2 = g->h * g->h;
    g->cell_size = 1.0f * g->h;
    g->kernel_factor = 0.5f;  // src/Simulation.hpp:147
    float h3 = (float)pow((double)g->h, 3.0);
    g->cubic_k = (float)(8.0f / (kPi * h3));
    g->cubic_l = (float)(48.0f / (kPi * h3));
    g->gX = (int)(g->domainX / g->cell_size) + 1;
    g->gY = (int)(g->domainY / g->cell_size) + 1;
    g->gZ = (int)(g->domainZ / g->cell_size) + 1;
    g->x_off = 0; g->slab = 0; g->x_lo = 0; g->x_hi = g->gX; g->gw = 0;
    g->gXg = g->gX; g->guard = cfg.slab_guard_columns > 0 ? cfg.slab_guard_columns : 0;
    if (cfg.slab_x_hi > cfg.slab_x_lo) {
        // spatial slab: local grid = owned columns [x_lo, x_hi) plus gw ghost columns on either side
        g->slab = 1;
        g->gw = cfg.slab_ghost_columns == 2 ? 2 : 1;
        g->x_lo = cfg.slab_x_lo; g->x_hi = cfg.slab_x_hi;
        g->x_off = cfg.slab_x_lo - g->gw;
        g->gX = cfg.slab_x_hi - cfg.slab_x_lo + 2 * g->gw;
    }
    g->gXZ = g->gX * g->gZ;
    g->C = g->gX * g->gY * g->gZ;
}

template <class T> static cudaError_t dalloc(T** p, size_t count) {
    *p = nullptr;
    return cudaMalloc((void**)p, sizeof(T) * (count ? count : 1));
}

int lgpu_scratch_reserve(lgpu_ctx* c, size_t bytes) {
    if (bytes <= c->scratch_bytes) return LGPU_OK;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (c->scratch) cudaFree(c->scratch);
    c->scratch = nullptr; c->scratch_bytes = 0;
    bytes = (bytes + (bytes >> 2) + 4095) & ~(size_t)4095;  // head room: the next request rarely reallocates
    CUDA_TRY(cudaMalloc((void**)&c->scratch, bytes));
    c->scratch_bytes = bytes;
    return LGPU_OK;
}

static int create_impl(const lgpu_config* cfg, lgpu_ctx* c, int device);
extern "C" int lgpu_create(const lgpu_config* cfg, lgpu_ctx** out) {
    if (!cfg || !out) return LGPU_ERR_ARG;
    *out = nullptr;
    if (cfg->domain[0] <= 0 || cfg->domain[1] <= 0 || cfg->domain[2] <= 0 || cfg->particle_radius <= 0.0f ||
        cfg->capacity_sand < 0 || cfg->capacity_solid < 0 || cfg->kernel_radius_scale <= 0.0f || cfg->slab_ghost_columns < 0 ||
        cfg->slab_ghost_columns > 2 || cfg->slab_guard_columns < 0 || (cfg->slab_x_hi > cfg->slab_x_lo && cfg->slab_x_hi - cfg->slab_x_lo < cfg->slab_ghost_columns)) {
        lgpu_set_error("lgpu_create: bad configuration");
        return LGPU_ERR_ARG;
    }
    int device = cfg->device;
    if (device < 0) CUDA_TRY(cudaGetDevice(&device));
    CUDA_TRY(cudaSetDevice(device));
    lgpu_ctx* c = new lgpu_ctx();
    memset(c, 0, sizeof(*c));
    const int st = create_impl(cfg, c, device);
    if (st != LGPU_OK) {  // nothing leaks: lgpu_destroy frees whatever was allocated so far (the context is zero-filled)
        char keep[sizeof(g_err)];
        memcpy(keep, g_err, sizeof(keep));
        lgpu_destroy(c);
        memcpy(g_err, keep, sizeof(keep));
        return st;
    }
    *out = c;
    return LGPU_OK;
}

static int create_impl(const lgpu_config* cfg, lgpu_ctx* c, int device) {
    c->cfg = *cfg;
    c->device = device;
    make_geom(*cfg, &c->g);
    c->cap = cfg->capacity_sand > 0 ? cfg->capacity_sand : 1;
    c->cap_solid = cfg->capacity_solid;
    c->M = cfg->max_neighbors > 0 ? cfg->max_neighbors : LGPU_DEFAULT_MAX_NEIGHBORS;
    c->M = (c->M + 3) & ~3;
    if (c->M > 4 * LGPU_MG) c->M = 4 * LGPU_MG;  // a row is at most LGPU_MG groups of four codes; longer lists re-walk the stencil
    c->stage_slots = LGPU_STAGE_SLOTS;
    // brick grid of the staged kernels (lgpu_neighbors.cuh); persistent kernels launch one block per SM
    c->nbY = (c->g.gY + LGPU_BY - 1) / LGPU_BY; c->nbX = (c->g.gX + LGPU_BX - 1) / LGPU_BX; c->nbZ = (c->g.gZ + LGPU_BZ - 1) / LGPU_BZ;
    c->NB = c->nbY * c->nbX * c->nbZ;
    CUDA_TRY(cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device));
    if (cfg->stream) { c->stream = (cudaStream_t)cfg->stream; c->own_stream = false; }
    else { CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
    const size_t cap = (size_t)c->cap, C1 = (size_t)c->g.C + 2;  // + trash cell (slab mode) + end
    for (int k = 0; k < 2; k++) {
        CUDA_TRY(dalloc(&c->pos[k], cap)); CUDA_TRY(dalloc(&c->vel[k], cap));
        CUDA_TRY(dalloc(&c->flags[k], cap)); CUDA_TRY(dalloc(&c->orig[k], cap));
    }
    CUDA_TRY(dalloc(&c->pstar_unsorted, cap));
    if (!c->g.slab) {  // slab mode: x0 / pa / pb live in the peer-visible arena (lgpu_slab_init)
        CUDA_TRY(dalloc(&c->x0, cap)); CUDA_TRY(dalloc(&c->pa, cap)); CUDA_TRY(dalloc(&c->pb, cap));
    }
    // every kernel of the step is loaded, and the staged ones opted into their dynamic shared memory, on THIS
    // device (both are per device; a process may hold contexts on several GPUs)
    if (lgpu_preload_grid() | lgpu_preload_neighbors() | lgpu_preload_fluid() | lgpu_preload_sand()) return LGPU_ERR_CUDA;
    if (c->g.slab) { int st = lgpu_slab_init(c); if (st) return st; }
    // key_in / rank_in / tmp_id double as scan input / output over the PARTICLES in lgpu_remove_in_cells and
    // lgpu_aabb_first_k: the scan writes starts[n] (one past the end) and one status word per 4096 elements
    CUDA_TRY(dalloc(&c->perm, cap)); CUDA_TRY(dalloc(&c->key_in, cap + 4)); CUDA_TRY(dalloc(&c->rank_in, cap + 4));
    CUDA_TRY(dalloc(&c->tmp_id, cap + 4)); CUDA_TRY(dalloc(&c->key, cap));
    CUDA_TRY(dalloc(&c->cell_count, C1)); CUDA_TRY(dalloc(&c->cell_start, C1));
    {   // scan state: three counters and one status word per 16384-element tile, zeroed once (the scan re-arms it itself)
        const size_t words = (C1 > cap + 4 ? C1 : cap + 4) / 4096 + 8;
        CUDA_TRY(dalloc(&c->scan_state, words));
        CUDA_TRY(cudaMemsetAsync(c->scan_state, 0, sizeof(unsigned long long) * words, c->stream));
    }
    CUDA_TRY(dalloc(&c->solid_pos, (size_t)c->cap_solid)); CUDA_TRY(dalloc(&c->solid_pos_unsorted, (size_t)c->cap_solid));
    CUDA_TRY(dalloc(&c->solid_orig, (size_t)c->cap_solid)); CUDA_TRY(dalloc(&c->solid_cell_start, C1));
    // table blocks: a meta word and LGPU_MG code words per own particle, rows padded to an even count per brick
    // (a non-empty brick part holds at least one particle; a brick has at most LGPU_BZ parts)
    const size_t nonempty = (size_t)((size_t)c->NB * LGPU_BZ < (size_t)c->cap ? (size_t)c->NB * LGPU_BZ : (size_t)c->cap);
    CUDA_TRY(dalloc(&c->nbr16, (cap + nonempty + 2) * (size_t)(1 + LGPU_MG))); CUDA_TRY(dalloc(&c->nbr_cnt, cap));
    // spill chunks for the tails of lists longer than the table width: one per four particles (or at least 1024)
    c->spill_cap = (int)(cap / 4 > 1024 ? cap / 4 : 1024);
    CUDA_TRY(dalloc(&c->nbr_spill, (size_t)c->spill_cap * (LGPU_SPILL / 4))); CUDA_TRY(dalloc(&c->nbr_ovf, cap));
    // a non-empty brick holds at least one particle, and there are at most NB bricks
    CUDA_TRY(dalloc(&c->brick_ctl, (size_t)(8 + LGPU_MAX_PASSES)));
    c->rec_cap = (int)nonempty + 1;
    CUDA_TRY(dalloc(&c->brick_rec, (size_t)c->rec_cap));
    CUDA_TRY(cudaMemsetAsync(c->brick_ctl, 0, sizeof(int) * (8 + LGPU_MAX_PASSES), c->stream));
    CUDA_TRY(dalloc(&c->lambda, cap)); CUDA_TRY(dalloc(&c->density, cap));
    CUDA_TRY(dalloc(&c->lambda_head, (size_t)LGPU_LAMBDA_HEAD));
    CUDA_TRY(dalloc(&c->counters, (size_t)4));
    c->stage_bytes = sizeof(float) * 8 * cap;  // pos 3 + vel 3 + flags 1 + ids 1 words per particle
    CUDA_TRY(cudaMalloc((void**)&c->d_stage, c->stage_bytes));
    CUDA_TRY(cudaMemsetAsync(c->cell_count, 0, sizeof(int) * C1, c->stream));
    CUDA_TRY(cudaMemsetAsync(c->cell_start, 0, sizeof(int) * C1, c->stream));
    CUDA_TRY(cudaMemsetAsync(c->solid_cell_start, 0, sizeof(int) * C1, c->stream));
    CUDA_TRY(cudaMemsetAsync(c->lambda, 0, sizeof(float) * cap, c->stream));
    CUDA_TRY(cudaMemsetAsync(c->lambda_head, 0, sizeof(float) * LGPU_LAMBDA_HEAD, c->stream));
    CUDA_TRY(cudaMemsetAsync(c->counters, 0, sizeof(unsigned long long) * 4, c->stream));
    CUDA_TRY(cudaMemsetAsync(c->nbr_cnt, 0, sizeof(int) * cap, c->stream));
    for (int k = 0; k < 2; k++) CUDA_TRY(cudaEventCreate(&c->ev[k]));
    for (int k = 0; k < LGPU_MAX_MARKS; k++) CUDA_TRY(cudaEventCreate(&c->ev_pool[k]));
    c->solids_sorted = true;  // no solids yet
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return LGPU_OK;
}

extern "C" void lgpu_destroy(lgpu_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->scratch);
    for (int k = 0; k < 2; k++) { cudaFree(c->pos[k]); cudaFree(c->vel[k]); cudaFree(c->flags[k]); cudaFree(c->orig[k]); }
    cudaFree(c->pstar_unsorted); cudaFree(c->perm);
    if (!c->g.slab) { cudaFree(c->x0); cudaFree(c->pa); cudaFree(c->pb); }
    cudaFree(c->key_in); cudaFree(c->rank_in); cudaFree(c->tmp_id); cudaFree(c->key);
    cudaFree(c->cell_count); cudaFree(c->cell_start); cudaFree(c->scan_state);
    cudaFree(c->solid_pos); cudaFree(c->solid_pos_unsorted); cudaFree(c->solid_orig); cudaFree(c->solid_cell_start);
    cudaFree(c->nbr16); cudaFree(c->nbr_cnt); cudaFree(c->nbr_spill); cudaFree(c->nbr_ovf); cudaFree(c->brick_ctl); cudaFree(c->brick_rec); cudaFree(c->lambda); cudaFree(c->density); cudaFree(c->lambda_head);
    cudaFree(c->counters); cudaFree(c->d_stage);
    lgpu_slab_free(c);
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    for (int k = 0; k < 2; k++) if (c->ev[k]) cudaEventDestroy(c->ev[k]);
    for (int k = 0; k < LGPU_MAX_MARKS; k++) if (c->ev_pool[k]) cudaEventDestroy(c->ev_pool[k]);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    cudaGetLastError();
    delete c;
}

View lgpu_make_view(lgpu_ctx* c) {
    View v;
    v.g = c->g;
    v.n = c->n; v.n_in = c->n_in; v.n_owned = c->n_owned; v.n_solid = c->n_solid; v.cap = c->cap; v.M = c->M;
    lgpu_slab_fill_view(c, &v);
    v.pos_in = c->pos[0]; v.vel_in = c->vel[0]; v.pstar_in = c->pstar_unsorted;
    v.flags_in = c->flags[0]; v.orig_in = c->orig[0];
    v.pos = c->pos[1]; v.vel = c->vel[1]; v.x0 = c->x0; v.pa = c->pa; v.pb = c->pb;
    v.flags = c->flags[1]; v.orig = c->orig[1]; v.perm = c->perm;
    v.key_in = c->key_in; v.rank_in = c->rank_in; v.tmp_id = c->tmp_id; v.key = c->key;
    v.sort_rec = reinterpret_cast<int4*>(c->vel[1]);
    v.cell_count = c->cell_count; v.cell_start = c->cell_start;
    v.solid_pos = c->solid_pos; v.solid_orig = c->solid_orig; v.solid_cell_start = c->solid_cell_start;
    v.nbr16 = c->nbr16; v.nbr_cnt = c->nbr_cnt; v.nbr_spill = c->nbr_spill; v.nbr_ovf = c->nbr_ovf; v.spill_cap = c->spill_cap;
    v.nbY = c->nbY; v.nbX = c->nbX; v.nbZ = c->nbZ; v.NB = c->NB; v.stage_slots = c->stage_slots;
    v.brick_ctl = c->brick_ctl; v.brick_rec = c->brick_rec; v.rec_cap = c->rec_cap;
    v.lambda = c->lambda; v.density = c->density; v.lambda_head = c->lambda_head;
    v.counters = c->counters;
    return v;
}

extern "C" int lgpu_get_grid(const lgpu_ctx* c, lgpu_grid_info* out) {
    if (!c || !out) return LGPU_ERR_ARG;
    out->grid[0] = c->g.gX; out->grid[1] = c->g.gY; out->grid[2] = c->g.gZ;
    out->num_cells = c->g.C;
    out->cell_size = c->g.cell_size; out->kernel_radius = c->g.h;
    out->cubic_k = c->g.cubic_k; out->cubic_l = c->g.cubic_l;
    return LGPU_OK;
}

extern "C" int lgpu_num_sand(const lgpu_ctx* c) { return c ? c->n_owned : 0; }
extern "C" int lgpu_num_solids(const lgpu_ctx* c) { return c ? c->n_solid : 0; }
extern "C" long lgpu_launch_count(const lgpu_ctx* c) { return c ? c->launches : 0; }
extern "C" int lgpu_set_phase_timing(lgpu_ctx* c, int on) { if (!c) return LGPU_ERR_ARG; c->phase_timing = on != 0; return LGPU_OK; }
extern "C" int lgpu_set_use_graph(lgpu_ctx* c, int on) {
    if (!c) return LGPU_ERR_ARG;
    if (on && c->g.slab) { lgpu_set_error("lgpu_set_use_graph: slab mode reads the migration counts on the host every substep; no graph"); return LGPU_ERR_ARG; }
    c->use_graph = on != 0;
    if (!on && c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
    return LGPU_OK;
}
extern "C" int lgpu_graph_stats(const lgpu_ctx* c, long* captures, long* replays) {
    if (!c) return LGPU_ERR_ARG;
    if (captures) *captures = c->graph_captures;
    if (replays) *replays = c->graph_replays;
    return LGPU_OK;
}
extern "C" int lgpu_set_stage_slots(lgpu_ctx* c, int slots) {
    if (!c || slots < 1) return LGPU_ERR_ARG;
    c->stage_slots = slots < LGPU_STAGE_SLOTS ? slots : LGPU_STAGE_SLOTS;
    if (c->stage_slots < LGPU_DUMMY_SLOTS) c->stage_slots = LGPU_DUMMY_SLOTS;
    return LGPU_OK;
}
extern "C" int lgpu_set_generic_kernels(lgpu_ctx* c, int on) {
    if (!c) return LGPU_ERR_ARG;
    c->generic_kernels = on != 0;
    return LGPU_OK;
}
extern "C" int lgpu_sync(lgpu_ctx* c) {
    if (!c) return LGPU_ERR_ARG;
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return lgpu_slab_check(c);
}

// ---------------- state transfer ----------------
__global__ void k_unpack3(const float* __restrict__ src, int n, float4* __restrict__ dst, int offset) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dst[offset + i] = make_float4(src[3 * i], src[3 * i + 1], src[3 * i + 2], 0.0f);
}
__global__ void k_fill_meta(int n, int offset, const int* __restrict__ flags_src, int* __restrict__ flags, int* __restrict__ orig,
                            float4* __restrict__ vel, int zero_vel, const int* __restrict__ ids) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flags[offset + i] = flags_src ? flags_src[i] : 0;
    orig[offset + i] = ids ? ids[i] : offset + i;
    if (zero_vel) vel[offset + i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}
// scatter to the reference slot: out[orig[i]] = src[i]
__global__ void k_pack3_by_orig(const float4* __restrict__ src, const int* __restrict__ orig, int n, float* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 a = src[i];
    int o = orig[i];
    dst[3 * o] = a.x; dst[3 * o + 1] = a.y; dst[3 * o + 2] = a.z;
}
__global__ void k_pack1_by_orig(const int* __restrict__ src, const int* __restrict__ orig, int n, int* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dst[orig[i]] = src[i];
}

// The device-side staging area holds all arrays of one transfer side by side (pos 3n, vel 3n, flags n,
// ids n: 8 words per particle), so a transfer is a run of back-to-back copies and kernels with ONE
// synchronisation at the end.
int lgpu_put_sand(lgpu_ctx* c, int offset, int n, const float* pos, const float* vel, const int* flags, const int* ids) {
    if (n == 0) return LGPU_OK;
    const int blocks = lgpu_blocks(n);
    float* d_pos = c->d_stage;
    float* d_vel = d_pos + 3 * (size_t)n;
    int* d_flags = (int*)(d_vel + 3 * (size_t)n);
    int* d_ids = d_flags + n;
    CUDA_TRY(cudaMemcpyAsync(d_pos, pos, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    if (vel) CUDA_TRY(cudaMemcpyAsync(d_vel, vel, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    if (flags) CUDA_TRY(cudaMemcpyAsync(d_flags, flags, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
    if (ids) CUDA_TRY(cudaMemcpyAsync(d_ids, ids, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
    k_unpack3<<<blocks, LGPU_BLOCK, 0, c->stream>>>(d_pos, n, c->pos[0], offset);
    if (vel) k_unpack3<<<blocks, LGPU_BLOCK, 0, c->stream>>>(d_vel, n, c->vel[0], offset);
    k_fill_meta<<<blocks, LGPU_BLOCK, 0, c->stream>>>(n, offset, flags ? d_flags : nullptr, c->flags[0], c->orig[0], c->vel[0], vel ? 0 : 1, ids ? d_ids : nullptr);
    c->launches += vel ? 3 : 2;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return LGPU_OK;
}

extern "C" int lgpu_upload_sand(lgpu_ctx* c, int n, const float* pos, const float* vel, const int* flags) {
    if (!c || n < 0 || (n > 0 && !pos)) return LGPU_ERR_ARG;
    if (n > c->cap) { lgpu_set_error("lgpu_upload_sand: %d particles > capacity %d", n, c->cap); return LGPU_ERR_CAPACITY; }
    CUDA_TRY(cudaSetDevice(c->device));
    c->n = c->n_owned = c->n_in = n;
    c->n_ghost = 0;
    c->grid_valid = false;
    CUDA_TRY(cudaMemsetAsync(c->lambda_head, 0, sizeof(float) * LGPU_LAMBDA_HEAD, c->stream));
    return lgpu_put_sand(c, 0, n, pos, vel, flags, nullptr);
}

extern "C" int lgpu_append_sand(lgpu_ctx* c, int n, const float* pos, const float* vel, const int* flags) {
    if (!c || n < 0 || (n > 0 && !pos)) return LGPU_ERR_ARG;
    if (c->n_owned + n > c->cap) { lgpu_set_error("lgpu_append_sand: capacity %d exceeded", c->cap); return LGPU_ERR_CAPACITY; }
    CUDA_TRY(cudaSetDevice(c->device));
    int st = lgpu_put_sand(c, c->n_owned, n, pos, vel, flags, nullptr);
    if (st) return st;
    c->n_owned += n;
    c->n = c->n_in = c->n_owned;
    c->grid_valid = false;
    return LGPU_OK;
}

extern "C" int lgpu_upload_solids(lgpu_ctx* c, int n, const float* pos) {
    if (!c || n < 0 || (n > 0 && !pos)) return LGPU_ERR_ARG;
    if (n > c->cap_solid) { lgpu_set_error("lgpu_upload_solids: %d > capacity %d", n, c->cap_solid); return LGPU_ERR_CAPACITY; }
    CUDA_TRY(cudaSetDevice(c->device));
    c->n_solid = c->n_solid_uploaded = n;
    c->solids_sorted = false;
    if (n > 0) {
        int st = lgpu_scratch_reserve(c, sizeof(float) * 3 * (size_t)n);
        if (st) return st;
        float* stage = (float*)c->scratch;
        CUDA_TRY(cudaMemcpyAsync(stage, pos, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c->stream));
        k_unpack3<<<lgpu_blocks(n), LGPU_BLOCK, 0, c->stream>>>(stage, n, c->solid_pos_unsorted, 0);
        c->launches++;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    return lgpu_sort_solids(c);
}

extern "C" int lgpu_download_sand2(lgpu_ctx* c, float* pos, float* pos2, float* vel, int* flags) {
    if (!c) return LGPU_ERR_ARG;
    CUDA_TRY(cudaSetDevice(c->device));
    const int n = c->n_owned;
    if (n == 0) return LGPU_OK;
    const int blocks = lgpu_blocks(n);
    float* d_pos = c->d_stage;
    float* d_vel = d_pos + 3 * (size_t)n;
    int* d_flags = (int*)(d_vel + 3 * (size_t)n);
    if (pos || pos2) {
        k_pack3_by_orig<<<blocks, LGPU_BLOCK, 0, c->stream>>>(c->pos[0], c->orig[0], n, d_pos);
        if (pos) CUDA_TRY(cudaMemcpyAsync(pos, d_pos, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
        c->launches++;
    }
    if (vel) {
        k_pack3_by_orig<<<blocks, LGPU_BLOCK, 0, c->stream>>>(c->vel[0], c->orig[0], n, d_vel);
        CUDA_TRY(cudaMemcpyAsync(vel, d_vel, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
        c->launches++;
    }
    if (flags) {
        k_pack1_by_orig<<<blocks, LGPU_BLOCK, 0, c->stream>>>(c->flags[0], c->orig[0], n, d_flags);
        CUDA_TRY(cudaMemcpyAsync(flags, d_flags, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
        c->launches++;
    }
    // (the second copy of the positions goes last: the caller's primary arrays are complete first)
    if (pos2) CUDA_TRY(cudaMemcpyAsync(pos2, d_pos, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return LGPU_OK;
}
extern "C" int lgpu_download_sand(lgpu_ctx* c, float* pos, float* vel, int* flags) { return lgpu_download_sand2(c, pos, nullptr, vel, flags); }

// ---------------- page-locked caller memory ----------------
extern "C" int lgpu_host_register(void* ptr, size_t bytes) {
    if (!ptr || !bytes) return LGPU_ERR_ARG;
    CUDA_TRY(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return LGPU_OK;
}
extern "C" int lgpu_host_unregister(void* ptr) {
    if (!ptr) return LGPU_ERR_ARG;
    CUDA_TRY(cudaHostUnregister(ptr));
    return LGPU_OK;
}
extern "C" int lgpu_host_is_pinned(const void* ptr) {
    if (!ptr) return 0;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) { cudaGetLastError(); return 0; }
    return a.type == cudaMemoryTypeHost ? 1 : 0;
}

// ---------------- step drivers ----------------
extern "C" int lgpu_slab_step_begin(lgpu_ctx* c, const lgpu_step_params* p, int mode);
extern "C" int lgpu_slab_step_end(lgpu_ctx* c);
int lgpu_begin_passes(lgpu_ctx* c) {
    // work-list counters and the table allocator of this substep
    c->pass = 0;
    CUDA_TRY(cudaMemsetAsync(c->brick_ctl, 0, sizeof(int) * (8 + LGPU_MAX_PASSES), c->stream));
    return LGPU_OK;
}

static int enqueue_step(lgpu_ctx* c, const lgpu_step_params& p, int mode) {
    int st;
    st = lgpu_begin_passes(c);
    if (st) return st;
    lgpu_mark(c, 1);
    st = mode == 1 ? lgpu_launch_predict_fluid(c, p) : lgpu_launch_predict_sand(c, p);
    if (st) return st;
    lgpu_mark(c, 2);
    st = lgpu_launch_scan_cells(c, c->cell_count, c->cell_start, c->g.C + 1, true, lgpu_pdl_enabled(c));  // + the trash cell of slab mode; chained to the predict kernel
    if (st) return st;
    lgpu_mark(c, 3);
    st = lgpu_launch_reorder(c, mode == 2);
    if (st) return st;
    lgpu_mark(c, 4);
    st = lgpu_launch_build_table(c, mode == 2, p, mode == 1 ? lgpu_fluid_lambda_mode(c, p) : LM_NONE);  // (fluid: + the first lambda pass)
    if (st) return st;
    st = mode == 1 ? lgpu_launch_fluid_solver(c, p) : lgpu_launch_sand_solver(c, p);  // marks 6 / 7 per launch
    if (st) return st;
    lgpu_mark(c, -1);
    return LGPU_OK;
}

// The substep as a CUDA graph: its launches are captured from the stream once (cudaStreamBeginCapture ..
// cudaStreamEndCapture around enqueue_step) and replayed with one cudaGraphLaunch while the particle count,
// the mode, the storage and the step parameters stay the same; when only parameters or counts change the
// instantiated graph is updated in place (cudaGraphExecUpdate), otherwise it is instantiated again.
static int step_as_graph(lgpu_ctx* c, const lgpu_step_params& p, int mode) {
    const long sig[5] = {mode, c->n, c->n_in, c->n_solid, (long)(size_t)c->pos[0]};
    if (c->graph_exec && memcmp(sig, c->graph_sig, sizeof(sig)) == 0 && memcmp(&p, &c->last_params, sizeof(p)) == 0) {
        CUDA_TRY(cudaGraphLaunch(c->graph_exec, c->stream));
        c->launches += c->graph_sig[5];
        c->graph_replays++;
        return LGPU_OK;
    }
    const long launches0 = c->launches;
    cudaGraph_t graph = nullptr;
    CUDA_TRY(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    const int st = enqueue_step(c, p, mode);
    const cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);  // always ended, also when a launch failed
    if (st) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return st; }
    if (ce != cudaSuccess) { lgpu_set_error("cudaStreamEndCapture -> %s", cudaGetErrorString(ce)); return LGPU_ERR_CUDA; }
    if (c->graph_exec) {
        cudaGraphExecUpdateResultInfo info;
        if (cudaGraphExecUpdate(c->graph_exec, graph, &info) != cudaSuccess) {
            cudaGetLastError();
            cudaGraphExecDestroy(c->graph_exec);
            c->graph_exec = nullptr;
        }
    }
    if (!c->graph_exec) {
        const cudaError_t ie = cudaGraphInstantiate(&c->graph_exec, graph, 0);
        if (ie != cudaSuccess) { cudaGraphDestroy(graph); c->graph_exec = nullptr; lgpu_set_error("cudaGraphInstantiate -> %s", cudaGetErrorString(ie)); return LGPU_ERR_CUDA; }
    }
    cudaGraphDestroy(graph);
    memcpy(c->graph_sig, sig, sizeof(sig));
    c->graph_sig[5] = c->launches - launches0;
    c->graph_captures++;
    CUDA_TRY(cudaGraphLaunch(c->graph_exec, c->stream));
    return LGPU_OK;
}

static int run_step(lgpu_ctx* c, const lgpu_step_params* p, int mode) {
    if (!c || !p) return LGPU_ERR_ARG;
    if (c->g.slab) {
        int st = lgpu_slab_step_begin(c, p, mode);
        return st ? st : lgpu_slab_step_end(c);
    }
    CUDA_TRY(cudaSetDevice(c->device));
    if (!c->solids_sorted) { int st = lgpu_sort_solids(c); if (st) return st; }
    c->last_mode = mode;
    c->n_marks = 0;
    CUDA_TRY(cudaEventRecord(c->ev[0], c->stream));
    // (per-launch event marks cannot be recorded inside a capture: phase timing runs eagerly)
    int st = (c->use_graph && !c->phase_timing) ? step_as_graph(c, *p, mode) : enqueue_step(c, *p, mode);
    c->last_params = *p;
    if (st) return st;
    CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
    c->grid_valid = true;
    return LGPU_OK;
}

extern "C" int lgpu_step_fluid(lgpu_ctx* c, const lgpu_step_params* p) { return run_step(c, p, 1); }
extern "C" int lgpu_step_sand(lgpu_ctx* c, const lgpu_step_params* p) { return run_step(c, p, 2); }

// Slab mode: a substep in two halves so that several contexts driven by ONE host thread ("virtual
// ranks", tests) can all post their halo messages before any of them waits.  One process per GPU
// simply calls lgpu_step_* (= begin + end).
extern "C" int lgpu_slab_step_begin(lgpu_ctx* c, const lgpu_step_params* p, int mode) {
    if (!c || !p || (mode != 1 && mode != 2) || !c->g.slab) return LGPU_ERR_ARG;
    CUDA_TRY(cudaSetDevice(c->device));
    if (!c->solids_sorted) { int st = lgpu_sort_solids(c); if (st) return st; }
    c->last_params = *p;
    c->last_mode = mode;
    c->n_marks = 0;
    CUDA_TRY(cudaEventRecord(c->ev[0], c->stream));
    return lgpu_slab_begin(c, *p, mode);
}
extern "C" int lgpu_slab_step_end(lgpu_ctx* c) {
    if (!c || !c->g.slab || !c->last_mode) return LGPU_ERR_ARG;
    CUDA_TRY(cudaSetDevice(c->device));
    int st = lgpu_slab_end(c, c->last_params, c->last_mode);
    if (st) return st;
    CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
    c->grid_valid = true;
    return LGPU_OK;
}

extern "C" int lgpu_last_step_ms(lgpu_ctx* c, int phase, float* ms) {
    if (!c || !ms || phase < 0 || phase > 8) return LGPU_ERR_ARG;
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaEventSynchronize(c->ev[1]));
    *ms = 0.0f;
    if (phase == 0) { CUDA_TRY(cudaEventElapsedTime(ms, c->ev[0], c->ev[1])); return LGPU_OK; }
    for (int k = 0; k + 1 < c->n_marks; k++) {
        int ph = c->ev_phase[k];
        if (ph == phase || (phase == 5 && (ph == 6 || ph == 7))) {
            float t = 0.0f;
            CUDA_TRY(cudaEventElapsedTime(&t, c->ev_pool[k], c->ev_pool[k + 1]));
            *ms += t;
        }
    }
    return LGPU_OK;
}

// ---------------- stage dumps ----------------
__global__ void k_dump_pstar(const float4* __restrict__ src, int n, float* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 a = src[i];
    dst[3 * i] = a.x; dst[3 * i + 1] = a.y; dst[3 * i + 2] = a.z;
}

__global__ void k_dump_cnt(const int* __restrict__ word, int n, int* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = word[i] & LGPU_CNT_MASK;
}

extern "C" int lgpu_dump(lgpu_ctx* c, int what, void* out, size_t out_bytes) {
    if (!c || !out) return LGPU_ERR_ARG;
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    const size_t n = (size_t)(c->g.slab ? c->n : c->n_owned);  // slab mode: the sorted live particles, ghosts included
    const void* src = nullptr;
    size_t bytes = 0;
    switch (what) {
        case LGPU_DUMP_KEYS: src = c->key; bytes = sizeof(int) * n; break;
        case LGPU_DUMP_PERM: src = c->perm; bytes = sizeof(int) * n; break;
        case LGPU_DUMP_ORIG: src = c->g.slab ? c->orig[1] : c->orig[0]; bytes = sizeof(int) * n; break;  // slab mode: particle ids, ghosts included
        case LGPU_DUMP_NBR_COUNT: {
            bytes = sizeof(int) * n;
            if (out_bytes < bytes) return LGPU_ERR_ARG;
            if (n == 0) return LGPU_OK;
            k_dump_cnt<<<lgpu_blocks((long)n), LGPU_BLOCK, 0, c->stream>>>(c->nbr_cnt, (int)n, (int*)c->d_stage);
            CUDA_TRY(cudaMemcpyAsync(out, c->d_stage, bytes, cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            return LGPU_OK;
        }
        case LGPU_DUMP_DENSITY: src = c->density; bytes = sizeof(float) * n; break;
        case LGPU_DUMP_LAMBDA: src = c->lambda; bytes = sizeof(float) * n; break;
        case LGPU_DUMP_CELL_START: src = c->cell_start; bytes = sizeof(int) * ((size_t)c->g.C + 1); break;
        case LGPU_DUMP_COUNTERS: src = c->counters; bytes = sizeof(unsigned long long) * 4; break;
        case LGPU_DUMP_PSTAR: {
            bytes = sizeof(float) * 3 * n;
            if (out_bytes < bytes) return LGPU_ERR_ARG;
            if (n == 0) return LGPU_OK;
            if (!c->pstar_final) return LGPU_ERR_ARG;
            k_dump_pstar<<<lgpu_blocks((long)n), LGPU_BLOCK, 0, c->stream>>>(c->pstar_final, (int)n, c->d_stage);
            CUDA_TRY(cudaMemcpyAsync(out, c->d_stage, bytes, cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            return LGPU_OK;
        }
        case LGPU_DUMP_NBR: {
            if (n == 0) return LGPU_OK;
            std::vector<int> cnt(n);
            CUDA_TRY(cudaMemcpy(cnt.data(), c->nbr_cnt, sizeof(int) * n, cudaMemcpyDeviceToHost));
            std::vector<long> off(n + 1);
            off[0] = 0;
            for (size_t i = 0; i < n; i++) off[i + 1] = off[i] + (cnt[i] & LGPU_CNT_MASK);
            bytes = sizeof(int) * (size_t)off[n];
            if (out_bytes < bytes) return LGPU_ERR_ARG;
            if (off[n] == 0) return LGPU_OK;
            long* d_off; int* d_flat;
            CUDA_TRY(cudaMalloc((void**)&d_off, sizeof(long) * (n + 1)));
            CUDA_TRY(cudaMalloc((void**)&d_flat, bytes));
            CUDA_TRY(cudaMemcpy(d_off, off.data(), sizeof(long) * (n + 1), cudaMemcpyHostToDevice));
            { int st = lgpu_launch_dump_nbr(c, c->last_mode == 2, d_off, d_flat); if (st) return st; }
            CUDA_TRY(cudaMemcpyAsync(out, d_flat, bytes, cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            cudaFree(d_off); cudaFree(d_flat);
            return LGPU_OK;
        }
        default: return LGPU_ERR_ARG;
    }
    if (out_bytes < bytes) { lgpu_set_error("lgpu_dump: buffer too small (%zu < %zu)", out_bytes, bytes); return LGPU_ERR_ARG; }
    if (bytes) CUDA_TRY(cudaMemcpy(out, src, bytes, cudaMemcpyDeviceToHost));
    return LGPU_OK;
}
