// lgpu_slab.cu — spatial slabs: one context per GPU owns the cell columns [x_lo, x_hi) of the
// reference's uniform grid.  The reference has no counterpart (it is single-threaded,
// src/Simulate.cpp); what is kept is its arithmetic: keys, neighbour sets and the per-particle
// updates of a slab run are those of the single-GPU run, because every owned particle sees the same
// neighbours (the one-cell ghost layer covers the support radius h = one cell).
//
// Per substep, between neighbouring slabs only (no collective):
//   1. after predict: particles whose predicted cell column left the slab MIGRATE (x, v, x*, flags,
//      id); particles in the first / last owned column are also sent as GHOST copies (k_push_halo);
//   2. after every solver pass but the last: the ghosts' x* (or just lambda) are REFRESHED from
//      their owners (k_refresh packs them into the neighbour's inbox, k_scatter_refresh unpacks).
// All traffic is written by the sender's kernels straight into the receiver's memory over
// NVLink (peer-mapped "arena", one CUDA IPC handle per context; a plain pointer for contexts that
// share a process), followed by a sequence-number flag that the receiver's kernels poll.  Both ends
// of a boundary enumerate the shared particles in message order, so no slot numbers travel.  Nothing
// goes through the host except four counters per substep.
#include <stdlib.h>
#include <string.h>

#include "lgpu_fluid.cuh"

struct SlabState {
    int halo_cap;
    size_t arena_bytes;
    unsigned char* arena;
    SlabArena local;
    SlabArena peer[2];
    void* peer_base[2];
    bool peer_ipc[2];
    int has_nbr[2];
    HaloRec *out_mig[2], *out_gho[2];
    int* gho_src[2];
    int* mig_src[2];
    int* out_cnt;         // 8 ints
    unsigned int* ticket; // 2 counters of the push kernels
    int* inv;
    int tx_seq[2], rx_seq[2];
    int n_gho_out[2], n_gho_in[2], n_mig_in[2], n_mig_out[2], in_base[2], mig_in_base[2];
    int* h_counts;        // pinned: [0..7] out_cnt, [8..] SlabHeader
    int* ref_src[2];      // ghost refresh, sending: sorted slot here of every particle that has a ghost copy in the neighbour
                          // on side s, in message order (ghost copies sent, then migrants received from that side)
    int* gho_slot[2];     // ghost refresh, receiving: sorted slot here of every ghost held from side s, same order as the
                          // neighbour's ref_src (ghost copies received, then the particles that emigrated to that side)
    int turn;             // which of the two refresh inboxes the next refresh uses
    int halo_turn;        // which of the two halo inbox sets this substep uses
    unsigned int* push_ticket;
    bool begun;
    int near_edge;        // particles inside the guard columns of an open side in the last substep (cropped plans)
    int n_store;
    void* d_xfer;         // device side of lgpu_slab_download (allocated once)
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// the same layout on the owner and on its neighbours
static size_t arena_layout(unsigned char* base, int cap, int halo_cap, SlabArena* a) {
    size_t off = 0;
    a->hdr = (SlabHeader*)(base + off); off += align_up(sizeof(SlabHeader), 256);
    for (int b = 0; b < 3; b++) { a->buf[b] = (float4*)(base + off); off += align_up(sizeof(float4) * (size_t)cap, 256); }
    for (int s = 0; s < 2; s++) for (int t = 0; t < 2; t++) { a->rbox[s][t] = (float4*)(base + off); off += align_up(sizeof(float4) * 2 * (size_t)halo_cap, 256); }
    for (int s = 0; s < 2; s++) for (int t = 0; t < 2; t++) { a->in_mig[s][t] = (HaloRec*)(base + off); off += align_up(sizeof(HaloRec) * (size_t)halo_cap, 256); }
    for (int s = 0; s < 2; s++) for (int t = 0; t < 2; t++) { a->in_gho[s][t] = (HaloRec*)(base + off); off += align_up(sizeof(HaloRec) * (size_t)halo_cap, 256); }
    return off;
}

int lgpu_preload_slab();
int lgpu_slab_init(lgpu_ctx* c) {
    SlabState* S = new SlabState();
    memset(S, 0, sizeof(*S));
    c->slab = S;
    S->halo_cap = c->cfg.halo_capacity > 0 ? c->cfg.halo_capacity : (c->cap / 4 > 65536 ? c->cap / 4 : 65536);
    SlabArena tmp;
    S->arena_bytes = arena_layout(nullptr, c->cap, S->halo_cap, &tmp);
    CUDA_TRY(cudaMalloc((void**)&S->arena, S->arena_bytes));
    CUDA_TRY(cudaMemsetAsync(S->arena, 0, align_up(sizeof(SlabHeader), 256), c->stream));
    arena_layout(S->arena, c->cap, S->halo_cap, &S->local);
    c->x0 = S->local.buf[0]; c->pa = S->local.buf[1]; c->pb = S->local.buf[2];
    for (int s = 0; s < 2; s++) {
        CUDA_TRY(cudaMalloc((void**)&S->out_mig[s], sizeof(HaloRec) * (size_t)S->halo_cap));
        CUDA_TRY(cudaMalloc((void**)&S->out_gho[s], sizeof(HaloRec) * (size_t)S->halo_cap));
        CUDA_TRY(cudaMalloc((void**)&S->gho_src[s], sizeof(int) * (size_t)S->halo_cap));
        CUDA_TRY(cudaMalloc((void**)&S->mig_src[s], sizeof(int) * (size_t)S->halo_cap));
    }
    CUDA_TRY(cudaMalloc((void**)&S->out_cnt, sizeof(int) * 8));
    CUDA_TRY(cudaMalloc((void**)&S->ticket, sizeof(unsigned int) * 2));
    CUDA_TRY(cudaMemsetAsync(S->ticket, 0, sizeof(unsigned int) * 2, c->stream));
    CUDA_TRY(cudaMalloc((void**)&S->inv, sizeof(int) * (size_t)c->cap));
    for (int s = 0; s < 2; s++) {
        CUDA_TRY(cudaMalloc((void**)&S->ref_src[s], sizeof(int) * 2 * (size_t)S->halo_cap));
        CUDA_TRY(cudaMalloc((void**)&S->gho_slot[s], sizeof(int) * 2 * (size_t)S->halo_cap));
    }
    CUDA_TRY(cudaMalloc((void**)&S->push_ticket, sizeof(unsigned int)));
    CUDA_TRY(cudaMemsetAsync(S->push_ticket, 0, sizeof(unsigned int), c->stream));
    CUDA_TRY(cudaMallocHost((void**)&S->h_counts, sizeof(int) * 64));
    // no kernel of the step may be loaded lazily while a neighbour waits for this context (see lgpu_grid.cu)
    int st = lgpu_preload_slab();  // (the kernels of the step itself were loaded by lgpu_create)
    return st ? LGPU_ERR_CUDA : LGPU_OK;
}

void lgpu_slab_free(lgpu_ctx* c) {
    SlabState* S = c->slab;
    if (!S) return;
    for (int s = 0; s < 2; s++) {
        if (S->peer_ipc[s] && S->peer_base[s]) cudaIpcCloseMemHandle(S->peer_base[s]);
        cudaFree(S->out_mig[s]); cudaFree(S->out_gho[s]); cudaFree(S->gho_src[s]); cudaFree(S->mig_src[s]);
    }
    cudaFree(S->arena); cudaFree(S->out_cnt); cudaFree(S->ticket); cudaFree(S->inv); cudaFree(S->ref_src[0]); cudaFree(S->ref_src[1]); cudaFree(S->gho_slot[0]); cudaFree(S->gho_slot[1]); cudaFree(S->push_ticket); cudaFree(S->d_xfer);
    cudaFreeHost(S->h_counts);
    delete S;
    c->slab = nullptr;
}

void lgpu_slab_fill_view(lgpu_ctx* c, View* v) {
    SlabState* S = c->slab;
    v->halo_cap = 0; v->has_nbr[0] = v->has_nbr[1] = 0; v->out_cnt = nullptr; v->inv = nullptr;
    for (int s = 0; s < 2; s++) { v->out_mig[s] = nullptr; v->out_gho[s] = nullptr; v->gho_src[s] = nullptr; v->mig_src[s] = nullptr; }
    if (!S) return;
    v->halo_cap = S->halo_cap; v->out_cnt = S->out_cnt; v->inv = S->inv;
    for (int s = 0; s < 2; s++) { v->has_nbr[s] = S->has_nbr[s]; v->out_mig[s] = S->out_mig[s]; v->out_gho[s] = S->out_gho[s]; v->gho_src[s] = S->gho_src[s]; v->mig_src[s] = S->mig_src[s]; }
}

// ---------------------------------------------------------------------------------------------
// C ABI: wiring
// ---------------------------------------------------------------------------------------------
extern "C" int lgpu_slab_export(lgpu_ctx* c, unsigned char handle[64], void** local_ptr, size_t* bytes) {
    if (!c || !c->slab) return LGPU_ERR_ARG;
    CUDA_TRY(cudaSetDevice(c->device));
    if (handle) {
        cudaIpcMemHandle_t h;
        CUDA_TRY(cudaIpcGetMemHandle(&h, c->slab->arena));
        static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
        memcpy(handle, &h, 64);
    }
    if (local_ptr) *local_ptr = c->slab->arena;
    if (bytes) *bytes = c->slab->arena_bytes;
    return LGPU_OK;
}

extern "C" int lgpu_slab_connect(lgpu_ctx* c, int side, const unsigned char handle[64], void* same_process_ptr) {
    if (!c || !c->slab || side < 0 || side > 1 || (!handle && !same_process_ptr)) return LGPU_ERR_ARG;
    CUDA_TRY(cudaSetDevice(c->device));
    SlabState* S = c->slab;
    void* base = same_process_ptr;
    if (!base) {
        cudaIpcMemHandle_t h;
        memcpy(&h, handle, 64);
        CUDA_TRY(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
        S->peer_ipc[side] = true;
    }
    S->peer_base[side] = base;
    arena_layout((unsigned char*)base, c->cap, S->halo_cap, &S->peer[side]);
    S->has_nbr[side] = 1;
    return LGPU_OK;
}

extern "C" int lgpu_slab_info(const lgpu_ctx* c, int out[8]) {
    if (!c || !out) return LGPU_ERR_ARG;
    out[0] = c->g.x_lo; out[1] = c->g.x_hi; out[2] = c->g.gX; out[3] = c->g.C;
    out[4] = c->n_owned; out[5] = c->n_ghost; out[6] = c->slab ? c->slab->halo_cap : 0; out[7] = c->g.x_off;
    return LGPU_OK;
}

extern "C" int lgpu_slab_edge(const lgpu_ctx* c, int out[2]) {
    if (!c || !out) return LGPU_ERR_ARG;
    const SlabState* S = c->slab;
    out[0] = S ? S->near_edge : 0;
    out[1] = S && ((!S->has_nbr[0] && c->g.x_lo > 0) || (!S->has_nbr[1] && c->g.x_hi < c->g.gXg)) ? 1 : 0;
    return LGPU_OK;
}

// ---------------------------------------------------------------------------------------------
// messages
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void signal_after_all_blocks(unsigned int* ticket, volatile int* flag, int seq) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1) {
            *ticket = 0;
            __threadfence_system();
            *flag = seq;
            __threadfence_system();
        }
    }
}

// halo message: migrants + ghost copies + their counts, then the flag
__global__ void __launch_bounds__(256) k_push_halo(const HaloRec* __restrict__ mig, const HaloRec* __restrict__ gho, const int* __restrict__ out_cnt, int side,
                                                   int halo_cap, HaloRec* __restrict__ dst_mig, HaloRec* __restrict__ dst_gho, SlabHeader* dst_hdr, int dst_side,
                                                   int turn, unsigned int* ticket, int seq) {
    const int nm = min(out_cnt[side], halo_cap), ng = min(out_cnt[2 + side], halo_cap);
    const uint4* s0 = reinterpret_cast<const uint4*>(mig);
    uint4* d0 = reinterpret_cast<uint4*>(dst_mig);
    const long wm = (long)nm * 4, wg = (long)ng * 4;  // 64-byte records = 4 x uint4
    for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < wm; t += (long)gridDim.x * blockDim.x) d0[t] = s0[t];
    const uint4* s1 = reinterpret_cast<const uint4*>(gho);
    uint4* d1 = reinterpret_cast<uint4*>(dst_gho);
    for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < wg; t += (long)gridDim.x * blockDim.x) d1[t] = s1[t];
    if (blockIdx.x == 0 && threadIdx.x == 0) { dst_hdr->in_cnt[turn][dst_side][0] = nm; dst_hdr->in_cnt[turn][dst_side][1] = ng; }
    signal_after_all_blocks(ticket, &dst_hdr->flag[dst_side], seq);
}

// Ghost-refresh lists, built locally once per substep after the sort.  Both ends of a slab boundary
// enumerate the shared particles in the same (message) order, so no slot numbers are exchanged:
//   receiving end: ghost g held from side s (ghost copies received, then the particles that emigrated
//                  to that side and stay here as ghosts) -> its sorted slot here;
//   sending end:   particle g mirrored on side s (ghost copies sent, then the migrants received from
//                  that side) -> its sorted slot here.
__global__ void __launch_bounds__(256) k_build_refresh_lists(const int* __restrict__ inv, int in_base, int n_gho_in, const int* __restrict__ mig_src, int n_mig_out,
                                                             int* __restrict__ gho_slot, const int* __restrict__ gho_src, int n_gho_out, int mig_in_base,
                                                             int n_mig_in, int* __restrict__ ref_src) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n_gho_in + n_mig_out) gho_slot[g] = inv[g < n_gho_in ? in_base + g : mig_src[g - n_gho_in]];
    if (g < n_gho_out + n_mig_in) ref_src[g] = inv[g < n_gho_out ? gho_src[g] : mig_in_base + (g - n_gho_out)];
}

// Ghost refresh after a solver pass, sending half: packs the value every mirrored particle got in the
// pass (x* as a float4, or only lambda as one float) CONTIGUOUSLY into the neighbour's refresh inbox —
// coalesced peer stores over NVLink; scattered 16- and 4-byte stores into the ghost slots ran at a
// fraction of the link rate —, then the last block raises the neighbours' sequence flags and waits
// for theirs.  The stores are NOT issued by the solver kernels themselves: a store to peer memory
// holds up the SM's memory pipeline for microseconds, which made the passes 20-45 % slower.
struct RefreshArgs {
    float4* buf;             // the buffer the pass wrote (read here, ghost slots written by k_scatter_refresh)
    int w_only;
    const int* src[2];       // ref_src
    int n_out[2];
    float4* peer_box[2];     // the neighbour's inbox for this slab's values
    volatile int* peer_flag[2];
    int seq[2];
    const volatile int* flag[2];
    int expected[2];
    unsigned int* ticket;
    int* error;
    const int* slot[2];      // gho_slot
    int n_in[2];
    const float4* box[2];    // this slab's inboxes
};
__global__ void __launch_bounds__(256) k_refresh(RefreshArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_trigger();  // (programmatic dependent launch: see pdl_trigger in lgpu_neighbors.cuh)
    int s = 0, g = 0, src = 0;
    if (t < a.n_out[0] + a.n_out[1]) {
        s = t < a.n_out[0] ? 0 : 1; g = t - (s ? a.n_out[0] : 0);
        src = a.src[s][g];
    }
    pdl_wait();     // the list above does not depend on the solver pass; the values do
    if (t < a.n_out[0] + a.n_out[1]) {
        const float4 val = a.buf[src];
        if (a.w_only) reinterpret_cast<float*>(a.peer_box[s])[g] = val.w;
        else a.peer_box[s][g] = val;
    }
    // one system-scope fence per block, by the thread that takes the ticket: the barrier makes the block's
    // stores visible to it, the fence orders them (cumulativity) before the ticket and the flags
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int ticket = atomicAdd(a.ticket, 1u);
        if (ticket == gridDim.x - 1) {
            *a.ticket = 0;
            __threadfence_system();
            if (a.peer_flag[0]) *a.peer_flag[0] = a.seq[0];
            if (a.peer_flag[1]) *a.peer_flag[1] = a.seq[1];
            const long long t0 = clock64();
            while ((a.flag[0] && *a.flag[0] < a.expected[0]) || (a.flag[1] && *a.flag[1] < a.expected[1])) {
                if (clock64() - t0 > 20000000000LL) { *a.error = 1; return; }  // ~10 s: the neighbour is gone; fail instead of hanging
                __nanosleep(40);
            }
            __threadfence_system();
        }
    }
}
// receiving half (next kernel on the stream): inbox -> ghost slots of the same buffer
__global__ void __launch_bounds__(256) k_scatter_refresh(RefreshArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_trigger();
    const bool live = t < a.n_in[0] + a.n_in[1];
    const int s = live && t >= a.n_in[0] ? 1 : 0, g = live ? t - (s ? a.n_in[0] : 0) : 0;
    const int slot = live ? a.slot[s][g] : 0;
    pdl_wait();     // k_refresh has seen the neighbours' flags: the inboxes are complete
    if (!live) return;
    if (a.w_only) {
        reinterpret_cast<float*>(a.buf + slot)[3] = __ldcv(reinterpret_cast<const float*>(a.box[s]) + g);
    } else {
        float4 val = __ldcv(a.box[s] + g);
        val.w = __int_as_float(slot);  // w of a full x* = the particle's slot where it is stored (sand solver)
        a.buf[slot] = val;
    }
}

__global__ void k_wait_flag(const volatile int* flag, int expected, int* error) {
    const long long t0 = clock64();
    while (*flag < expected) {
        if (clock64() - t0 > 20000000000LL) { *error = 1; return; }  // ~10 s: the neighbour is gone; fail instead of hanging
        __nanosleep(200);
    }
    __threadfence_system();
}

// inbox -> unsorted storage [n_store, n_store + total), keys + histogram
__global__ void __launch_bounds__(LGPU_BLOCK) k_append_halo(View v, int n_store, int m0, int m1, int g0, int g1, const HaloRec* __restrict__ im0,
                                                            const HaloRec* __restrict__ im1, const HaloRec* __restrict__ ig0, const HaloRec* __restrict__ ig1) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = m0 + m1 + g0 + g1;
    if (t >= total) return;
    const HaloRec* r;
    int ghost = 0;
    if (t < m0) r = im0 + t;
    else if (t < m0 + m1) r = im1 + (t - m0);
    else if (t < m0 + m1 + g0) { r = ig0 + (t - m0 - m1); ghost = 1; }
    else { r = ig1 + (t - m0 - m1 - g0); ghost = 1; }
    const HaloRec rec = *r;
    const int i = n_store + t;
    v.pos_in[i] = rec.pos; v.vel_in[i] = rec.vel; v.pstar_in[i] = rec.pstar;
    v.flags_in[i] = (rec.flags & ~(LGPU_FLAG_DEAD | LGPU_FLAG_GHOST)) | (ghost ? LGPU_FLAG_GHOST : 0);
    v.orig_in[i] = rec.orig;
    int key = cell_id_checked(v.g, f3(rec.pstar), v.counters);
    v.key_in[i] = key;
    v.rank_in[i] = atomicAdd(&v.cell_count[key], 1);
}

#define LGPU_PRELOAD(f) do { cudaFuncAttributes a; CUDA_TRY(cudaFuncGetAttributes(&a, f)); } while (0)
__global__ void k_compact_owned(View v, int n_store, float* __restrict__ pos, float* __restrict__ vel, int* __restrict__ flags, int* __restrict__ ids,
                                int* __restrict__ counter);
int lgpu_preload_slab() {
    LGPU_PRELOAD(k_push_halo); LGPU_PRELOAD(k_build_refresh_lists); LGPU_PRELOAD(k_wait_flag); LGPU_PRELOAD(k_refresh); LGPU_PRELOAD(k_scatter_refresh);
    LGPU_PRELOAD(k_append_halo); LGPU_PRELOAD(k_compact_owned);
    return LGPU_OK;
}

// a wait kernel that timed out leaves its mark in the arena header
int lgpu_slab_check(lgpu_ctx* c) {
    SlabState* S = c->slab;
    if (!S) return LGPU_OK;
    int err = 0;
    CUDA_TRY(cudaMemcpyAsync(&err, &S->local.hdr->error, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (err) { lgpu_set_error("slab: timed out waiting for a message of a neighbouring slab"); return LGPU_ERR_CUDA; }
    return LGPU_OK;
}

// ---------------------------------------------------------------------------------------------
// step halves
// ---------------------------------------------------------------------------------------------
int lgpu_slab_begin(lgpu_ctx* c, const lgpu_step_params& p, int mode) {
    SlabState* S = c->slab;
    if (mode == 1 && p.literal_lambda_index && (S->has_nbr[0] || S->has_nbr[1])) {
        lgpu_set_error("slab mode: literal_lambda_index=1 (lambdas[loop counter], SURVEY F4) depends on the global particle numbering and is "
                       "only defined on a single GPU; use literal_lambda_index=0");
        return LGPU_ERR_ARG;
    }
    S->n_store = c->n_in = c->n;  // the step-boundary storage: last step's sorted slots (ghost slots are dead)
    { int st0 = lgpu_begin_passes(c); if (st0) return st0; }
    CUDA_TRY(cudaMemsetAsync(S->out_cnt, 0, sizeof(int) * 8, c->stream));
    lgpu_mark(c, 1);
    int st = mode == 1 ? lgpu_launch_predict_fluid(c, p) : lgpu_launch_predict_sand(c, p);
    if (st) return st;
    for (int s = 0; s < 2; s++) {
        if (!S->has_nbr[s]) continue;
        const int ds = 1 - s;  // my right neighbour receives from its left
        k_push_halo<<<256, 256, 0, c->stream>>>(S->out_mig[s], S->out_gho[s], S->out_cnt, s, S->halo_cap, S->peer[s].in_mig[ds][S->halo_turn],
                                               S->peer[s].in_gho[ds][S->halo_turn], S->peer[s].hdr, ds, S->halo_turn, S->ticket + s, ++S->tx_seq[s]);
        c->launches++;
    }
    CUDA_TRY(cudaGetLastError());
    S->begun = true;
    return LGPU_OK;
}

int lgpu_slab_end(lgpu_ctx* c, const lgpu_step_params& p, int mode) {
    SlabState* S = c->slab;
    if (!S->begun) return LGPU_ERR_ARG;
    S->begun = false;
    for (int s = 0; s < 2; s++) {
        if (!S->has_nbr[s]) continue;
        k_wait_flag<<<1, 1, 0, c->stream>>>(&S->local.hdr->flag[s], ++S->rx_seq[s], &S->local.hdr->error);
        c->launches++;
    }
    // the only host round trip of the substep: how many particles left, arrived, are ghosts
    CUDA_TRY(cudaMemcpyAsync(S->h_counts, S->out_cnt, sizeof(int) * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(S->h_counts + 8, S->local.hdr, sizeof(SlabHeader), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    const SlabHeader* H = (const SlabHeader*)(S->h_counts + 8);
    if (H->error) { lgpu_set_error("slab: timed out waiting for a neighbouring slab's halo message"); return LGPU_ERR_CUDA; }
    if (S->h_counts[4]) { lgpu_set_error("slab: halo capacity %d exceeded (%d records dropped)", S->halo_cap, S->h_counts[4]); return LGPU_ERR_CAPACITY; }
    S->near_edge = S->h_counts[5];
    if (S->h_counts[6]) {
        lgpu_set_error("slab: %d particle(s) left the planned cell columns [%d, %d) on a side that has no neighbouring slab (a plan cropped to the "
                       "occupied columns): re-plan earlier (lgpu_slab_edge / replan_if_needed) or plan with a wider margin", S->h_counts[6], c->g.x_lo, c->g.x_hi);
        return LGPU_ERR_CAPACITY;
    }
    int m[2] = {0, 0}, g[2] = {0, 0};
    const int ht = S->halo_turn;
    S->halo_turn ^= 1;
    for (int s = 0; s < 2; s++) if (S->has_nbr[s]) { m[s] = H->in_cnt[ht][s][0]; g[s] = H->in_cnt[ht][s][1]; }
    const int emigrated = (S->has_nbr[0] ? S->h_counts[0] : 0) + (S->has_nbr[1] ? S->h_counts[1] : 0);
    const int in_total = m[0] + m[1] + g[0] + g[1];
    if (S->n_store + in_total > c->cap) { lgpu_set_error("slab: %d + %d particles > capacity %d", S->n_store, in_total, c->cap); return LGPU_ERR_CAPACITY; }
    for (int s = 0; s < 2; s++) {
        S->n_gho_out[s] = S->has_nbr[s] ? S->h_counts[2 + s] : 0; S->n_mig_out[s] = S->has_nbr[s] ? S->h_counts[s] : 0;
        S->n_gho_in[s] = g[s]; S->n_mig_in[s] = m[s];
    }
    S->mig_in_base[0] = S->n_store;
    S->mig_in_base[1] = S->n_store + m[0];
    S->in_base[0] = S->n_store + m[0] + m[1];
    S->in_base[1] = S->in_base[0] + g[0];
    const int dead_before = c->n_ghost;  // last step's ghost slots
    c->n_in = S->n_store + in_total;
    c->n = S->n_store - dead_before + in_total;  // emigrants stay as ghosts of their new owner
    c->n_ghost = g[0] + g[1] + emigrated;
    c->n_owned = c->n - c->n_ghost;
    if (in_total > 0) {
        View v = lgpu_make_view(c);
        k_append_halo<<<lgpu_blocks(in_total), LGPU_BLOCK, 0, c->stream>>>(v, S->n_store, m[0], m[1], g[0], g[1], S->local.in_mig[0][ht], S->local.in_mig[1][ht],
                                                                            S->local.in_gho[0][ht], S->local.in_gho[1][ht]);
        c->launches++;
    }
    int st;
    lgpu_mark(c, 2);
    st = lgpu_launch_scan_cells(c, c->cell_count, c->cell_start, c->g.C + 1, true, false);
    if (st) return st;
    lgpu_mark(c, 3);
    st = lgpu_launch_reorder(c, false);  // particle ids are global: never renumbered
    if (st) return st;
    // ghost-refresh lists of this substep (local: both ends enumerate the shared particles in message order)
    for (int s = 0; s < 2; s++) {
        if (!S->has_nbr[s]) continue;
        const int n_in = S->n_gho_in[s] + S->n_mig_out[s], n_out = S->n_gho_out[s] + S->n_mig_in[s];
        const int n = n_in > n_out ? n_in : n_out;
        if (n > 0) {
            k_build_refresh_lists<<<(n + 255) / 256, 256, 0, c->stream>>>(S->inv, S->in_base[s], S->n_gho_in[s], S->mig_src[s], S->n_mig_out[s], S->gho_slot[s],
                                                                         S->gho_src[s], S->n_gho_out[s], S->mig_in_base[s], S->n_mig_in[s], S->ref_src[s]);
            c->launches++;
        }
    }
    lgpu_mark(c, 4);
    st = lgpu_launch_build_table(c, mode == 2, p, mode == 1 ? lgpu_fluid_lambda_mode(c, p) : LM_NONE);
    if (st) return st;
    st = mode == 1 ? lgpu_launch_fluid_solver(c, p) : lgpu_launch_sand_solver(c, p);
    if (st) return st;
    lgpu_mark(c, -1);
    CUDA_TRY(cudaGetLastError());
    return LGPU_OK;
}

// Programmatic dependent launch keeps the next solver kernel's blocks resident while k_refresh waits for the
// neighbours.  That is only safe when every neighbour runs on ANOTHER GPU: slabs that share a device (the
// virtual ranks of the single-GPU tests) would starve each other of SM resources and dead-lock.
// Measured: 1.5 % per substep on 2 B200s, 2 % on 8 (1.27 -> 1.24 ms, 16 M dam break); bit-identical to the single-context
// run.  On whenever every neighbour is another process's GPU; LGPU_PDL_SLAB=0 turns it off.
bool lgpu_slab_pdl_ok(const lgpu_ctx* c) {
    const SlabState* S = c->slab;
    if (!S || (!S->has_nbr[0] && !S->has_nbr[1])) return true;
    static const bool off = getenv("LGPU_PDL_SLAB") && atoi(getenv("LGPU_PDL_SLAB")) == 0;
    if (off) return false;
    for (int s = 0; s < 2; s++) if (S->has_nbr[s] && !S->peer_ipc[s]) return false;
    return true;
}

bool lgpu_slab_active(const lgpu_ctx* c) {
    const SlabState* S = c->slab;
    return S && (S->has_nbr[0] || S->has_nbr[1]);
}

// after a solver kernel that wrote `buf`: send the mirrored particles' new values to the neighbours' inboxes,
// signal, wait (k_refresh), then move what the neighbours sent into this slab's ghost slots (k_scatter_refresh)
int lgpu_slab_refresh(lgpu_ctx* c, const float4* buf, bool w_only) {
    SlabState* S = c->slab;
    if (!lgpu_slab_active(c)) return LGPU_OK;
    const int turn = S->turn;
    S->turn ^= 1;
    RefreshArgs a;
    memset(&a, 0, sizeof(a));
    a.buf = const_cast<float4*>(buf);
    a.w_only = w_only ? 1 : 0;
    a.ticket = S->push_ticket;
    a.error = &S->local.hdr->error;
    for (int s = 0; s < 2; s++) {
        if (!S->has_nbr[s]) continue;
        a.src[s] = S->ref_src[s];
        a.n_out[s] = S->n_gho_out[s] + S->n_mig_in[s];
        a.peer_box[s] = S->peer[s].rbox[1 - s][turn];
        a.peer_flag[s] = &S->peer[s].hdr->flag[1 - s];
        a.seq[s] = ++S->tx_seq[s];
        a.flag[s] = &S->local.hdr->flag[s];
        a.expected[s] = ++S->rx_seq[s];
        a.slot[s] = S->gho_slot[s];
        a.n_in[s] = S->n_gho_in[s] + S->n_mig_out[s];
        a.box[s] = S->local.rbox[s][turn];
    }
    const int n_out = a.n_out[0] + a.n_out[1], n_in = a.n_in[0] + a.n_in[1];
    static const bool pdl_env = !(getenv("LGPU_PDL") && atoi(getenv("LGPU_PDL")) == 0);
    const bool pdl = pdl_env && !c->phase_timing && !c->use_graph && lgpu_slab_pdl_ok(c);
    CUDA_TRY(launch_pdl(k_refresh, n_out > 0 ? (n_out + 255) / 256 : 1, 256, 0, c->stream, pdl, a));
    c->launches++;
    if (n_in > 0) {
        CUDA_TRY(launch_pdl(k_scatter_refresh, (n_in + 255) / 256, 256, 0, c->stream, pdl, a));
        c->launches++;
    }
    CUDA_TRY(cudaGetLastError());
    return LGPU_OK;
}

// ---------------------------------------------------------------------------------------------
// state transfer in slab mode: particles carry caller-given global ids
// ---------------------------------------------------------------------------------------------
extern "C" int lgpu_slab_upload(lgpu_ctx* c, int n, const float* pos, const float* vel, const int* flags, const int* ids) {
    if (!c || !c->slab || n < 0 || (n > 0 && (!pos || !ids))) return LGPU_ERR_ARG;
    if (n > c->cap) { lgpu_set_error("lgpu_slab_upload: %d particles > capacity %d", n, c->cap); return LGPU_ERR_CAPACITY; }
    CUDA_TRY(cudaSetDevice(c->device));
    c->n = c->n_owned = c->n_in = n;
    c->n_ghost = 0;
    c->grid_valid = false;
    c->slab->near_edge = 0;
    return lgpu_put_sand(c, 0, n, pos, vel, flags, ids);
}

__global__ void __launch_bounds__(LGPU_BLOCK) k_compact_owned(View v, int n_store, float* __restrict__ pos, float* __restrict__ vel, int* __restrict__ flags,
                                                              int* __restrict__ ids, int* __restrict__ counter) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_store) return;
    const int a = v.flags_in[i];
    const bool live = !(a & (LGPU_FLAG_DEAD | LGPU_FLAG_GHOST));
    const int o = warp_agg_inc(counter, live);  // (one atomic per warp: millions of lanes on one counter serialise)
    if (!live) return;
    const float4 x = v.pos_in[i], u = v.vel_in[i];
    pos[3 * o] = x.x; pos[3 * o + 1] = x.y; pos[3 * o + 2] = x.z;
    vel[3 * o] = u.x; vel[3 * o + 1] = u.y; vel[3 * o + 2] = u.z;
    flags[o] = a; ids[o] = v.orig_in[i];
}

// Owned particles of this slab in storage order is not defined across slabs: the caller gets
// (id, x, v, flags) tuples and orders them by id.  All output buffers hold lgpu_num_sand() entries.
extern "C" int lgpu_slab_download(lgpu_ctx* c, float* pos, float* vel, int* flags, int* ids, int* n_out) {
    if (!c || !c->slab || !pos || !vel || !flags || !ids || !n_out) return LGPU_ERR_ARG;
    CUDA_TRY(cudaSetDevice(c->device));
    const int n_store = c->n, n = c->n_owned;
    *n_out = 0;
    if (n == 0) return LGPU_OK;
    // one transfer area per context, allocated at the first download and kept: cudaMalloc / cudaFree
    // are device-wide synchronisation points and, with peer mappings in place, cost milliseconds
    SlabState* S = c->slab;
    const size_t need = sizeof(int) * 8 * (size_t)c->cap + 256;
    if (!S->d_xfer) { CUDA_TRY(cudaMalloc((void**)&S->d_xfer, need)); }
    float* d_pos = (float*)S->d_xfer;
    float* d_vel = d_pos + 3 * (size_t)c->cap;
    int* d_flags = (int*)(d_vel + 3 * (size_t)c->cap);
    int* d_ids = d_flags + c->cap;
    int* d_counter = d_ids + c->cap;
    CUDA_TRY(cudaMemsetAsync(d_counter, 0, sizeof(int), c->stream));
    View v = lgpu_make_view(c);
    k_compact_owned<<<lgpu_blocks(n_store), LGPU_BLOCK, 0, c->stream>>>(v, n_store, d_pos, d_vel, d_flags, d_ids, d_counter);
    c->launches++;
    int got = 0;
    CUDA_TRY(cudaMemcpyAsync(&got, d_counter, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(pos, d_pos, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(vel, d_vel, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(flags, d_flags, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(ids, d_ids, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (got != n) { lgpu_set_error("lgpu_slab_download: %d owned particles found, %d expected", got, n); return LGPU_ERR_CUDA; }
    *n_out = n;
    return LGPU_OK;
}
