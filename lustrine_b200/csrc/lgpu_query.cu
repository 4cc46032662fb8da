// lgpu_query.cu — the callers either side of the step that index the uniform grid on the host
// in the reference ("next" rows): query_cell_num_particles (src/Lustrine.cpp:864-914),
// the sink pass of Lustrine::simulate (:806-836) and the player-AABB scan that feeds the Bullet
// proxy boxes (src/BulletPhysics.cpp:602-652).  All of them work on the device grid / storage of
// the last step; only a few hundred bytes cross PCIe.
#include <string.h>

#include <unordered_map>
#include <vector>

#include "lgpu_internal.cuh"

// ---------------- query_cell_num_particles ----------------
__global__ void k_cell_count(View v, int x0, int y0, int z0, int nx, int ny, int nz, int include_solid, unsigned long long* out) {
    long total = (long)nx * ny * nz;
    unsigned long long acc = 0;
    for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
        int z = (int)(t % nz);
        int x = (int)((t / nz) % nx);
        int y = (int)(t / ((long)nz * nx));
        int c = (y0 + y) * v.g.gXZ + (x0 + x) * v.g.gZ + (z0 + z);
        acc += (unsigned long long)(v.cell_start[c + 1] - v.cell_start[c]);
        if (include_solid) acc += (unsigned long long)(v.solid_cell_start[c + 1] - v.solid_cell_start[c]);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

extern "C" int lgpu_cell_count(lgpu_ctx* c, const int lo[3], const int hi[3], int include_solid, int* count) {
    if (!c || !lo || !hi || !count) return LGPU_ERR_ARG;
    CUDA_TRY(cudaSetDevice(c->device));
    *count = 0;
    int x0 = lo[0] < 0 ? 0 : lo[0], y0 = lo[1] < 0 ? 0 : lo[1], z0 = lo[2] < 0 ? 0 : lo[2];
    int x1 = hi[0] >= c->g.gX ? c->g.gX - 1 : hi[0], y1 = hi[1] >= c->g.gY ? c->g.gY - 1 : hi[1], z1 = hi[2] >= c->g.gZ ? c->g.gZ - 1 : hi[2];
    if (x1 < x0 || y1 < y0 || z1 < z0) return LGPU_OK;
    View v = lgpu_make_view(c);
    unsigned long long* d_out = c->counters + 2;
    CUDA_TRY(cudaMemsetAsync(d_out, 0, sizeof(unsigned long long), c->stream));
    long total = (long)(x1 - x0 + 1) * (y1 - y0 + 1) * (z1 - z0 + 1);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 1184) blocks = 1184;
    k_cell_count<<<blocks, 256, 0, c->stream>>>(v, x0, y0, z0, x1 - x0 + 1, y1 - y0 + 1, z1 - z0 + 1, include_solid, d_out);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    unsigned long long h = 0;
    CUDA_TRY(cudaMemcpyAsync(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *count = (int)h;
    return LGPU_OK;
}

// ---------------- sinks ----------------
__global__ void k_sink_counts(View v, const int* __restrict__ cells, int n_cells, int* __restrict__ counts) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_cells) return;
    int c = cells[t];
    counts[t] = (c >= 0 && c < v.g.C) ? v.cell_start[c + 1] - v.cell_start[c] : 0;
}
// evicted reference slots in the reference's visiting order: sink cells as given, particles of a
// cell in ascending reference slot (= ascending sorted slot)
__global__ void k_sink_list(View v, const int* __restrict__ cells, const int* __restrict__ offsets, int n_cells, int* __restrict__ list) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_cells) return;
    int c = cells[t];
    if (c < 0 || c >= v.g.C) return;
    int b = v.cell_start[c], e = v.cell_start[c + 1], o = offsets[t];
    for (int u = b; u < e; u++) list[o + (u - b)] = v.orig_in[u];
}
__global__ void k_invert_orig(const int* __restrict__ orig, int n, int* __restrict__ inv) {
    int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u < n) inv[orig[u]] = u;
}
__global__ void k_apply_moves(const int* __restrict__ inv, const int* __restrict__ dead, int n_dead, const int* __restrict__ moved, int n_moved,
                              int* __restrict__ keep, int* __restrict__ orig) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_dead) keep[inv[dead[t]]] = 0;
    if (t < n_moved) orig[inv[moved[2 * t]]] = moved[2 * t + 1];
}
__global__ void k_fill_int(int* p, int n, int value) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = value;
}
__global__ void k_compact(View v, const int* __restrict__ keep, const int* __restrict__ dst_index, int n) {
    int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n || !keep[u]) return;
    int d = dst_index[u];
    v.pos[d] = v.pos_in[u];
    v.vel[d] = v.vel_in[u];
    v.flags[d] = v.flags_in[u];
    v.orig[d] = v.orig_in[u];
}

extern "C" int lgpu_remove_in_cells(lgpu_ctx* c, const int* cell_ids, int n_cells, int* removed) {
    if (!c || (n_cells > 0 && !cell_ids) || n_cells < 0) return LGPU_ERR_ARG;
    if (removed) *removed = 0;
    if (n_cells == 0 || c->n_owned == 0) return LGPU_OK;
    if (!c->grid_valid) { lgpu_set_error("lgpu_remove_in_cells: no grid (call a step first)"); return LGPU_ERR_ARG; }
    CUDA_TRY(cudaSetDevice(c->device));
    View v = lgpu_make_view(c);
    const int n = c->n_owned;
    // all temporaries of this call come from the context's scratch buffer (no per-frame cudaMalloc / cudaFree):
    // sink cells, their counts / offsets, the evicted slots (<= n), dead slots (<= n), moves (<= n pairs)
    const size_t nc4 = ((size_t)n_cells + 3) & ~(size_t)3, n4 = ((size_t)n + 3) & ~(size_t)3;
    { int st = lgpu_scratch_reserve(c, sizeof(int) * (2 * nc4 + 4 * n4)); if (st) return st; }
    int* d_cells = (int*)c->scratch;
    int* d_counts = d_cells + nc4;
    int* d_list = d_counts + nc4;
    int* d_dead = d_list + n4;
    int* d_moved = d_dead + n4;
    CUDA_TRY(cudaMemcpyAsync(d_cells, cell_ids, sizeof(int) * n_cells, cudaMemcpyHostToDevice, c->stream));
    k_sink_counts<<<lgpu_blocks(n_cells), LGPU_BLOCK, 0, c->stream>>>(v, d_cells, n_cells, d_counts);
    std::vector<int> counts(n_cells), offsets(n_cells);
    CUDA_TRY(cudaMemcpyAsync(counts.data(), d_counts, sizeof(int) * n_cells, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    int total = 0;
    for (int t = 0; t < n_cells; t++) { offsets[t] = total; total += counts[t]; }
    c->launches++;
    if (total == 0) return LGPU_OK;
    CUDA_TRY(cudaMemcpyAsync(d_counts, offsets.data(), sizeof(int) * n_cells, cudaMemcpyHostToDevice, c->stream));
    k_sink_list<<<lgpu_blocks(n_cells), LGPU_BLOCK, 0, c->stream>>>(v, d_cells, d_counts, n_cells, d_list);
    std::vector<int> list(total);
    CUDA_TRY(cudaMemcpyAsync(list.data(), d_list, sizeof(int) * total, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->launches++;
    // The reference's swap-with-last loop (src/Lustrine.cpp:825-834), replayed on slot numbers only:
    // content[s] = reference slot (before the pass) of the particle that ends up in slot s.
    std::unordered_map<int, int> content;
    auto get = [&](int s) { auto it = content.find(s); return it == content.end() ? s : it->second; };
    int end = n;
    for (int t = 0; t < total; t++) {
        int e = list[t];
        content[e] = get(end - 1);
        end--;
    }
    // Live slots are [0, end).  A particle (named by its old slot) survives iff it is the content
    // of a live slot; the copy source of every move dies at once, so no particle is duplicated.
    std::unordered_map<int, int> new_slot_of;  // old slot -> new slot, for particles that moved
    std::vector<int> dead;
    {
        std::unordered_map<int, int> seen;  // old slot -> live slot now holding that particle
        for (auto& kv : content) if (kv.first < end) seen[kv.second] = kv.first;
        std::vector<int> candidates;
        for (auto& kv : content) { candidates.push_back(kv.first); candidates.push_back(kv.second); }
        for (int s = end; s < n; s++) candidates.push_back(s);
        std::unordered_map<int, char> done;
        for (int old : candidates) {
            if (old < 0 || old >= n || done.count(old)) continue;
            done[old] = 1;
            int ns = -1;
            auto it = seen.find(old);
            if (it != seen.end()) ns = it->second;
            else if (old < end && content.find(old) == content.end()) ns = old;  // untouched
            if (ns < 0) dead.push_back(old);
            else if (ns != old) new_slot_of[old] = ns;
        }
    }
    std::vector<int> moved;
    for (auto& kv : new_slot_of) { moved.push_back(kv.first); moved.push_back(kv.second); }
    const int n_dead = (int)dead.size(), n_moved = (int)moved.size() / 2;
    int* inv = c->tmp_id;      // scratch, dead between steps
    int* keep = c->key_in;
    int* dst = c->rank_in;
    if (n_dead) CUDA_TRY(cudaMemcpyAsync(d_dead, dead.data(), sizeof(int) * n_dead, cudaMemcpyHostToDevice, c->stream));
    if (n_moved) CUDA_TRY(cudaMemcpyAsync(d_moved, moved.data(), sizeof(int) * 2 * n_moved, cudaMemcpyHostToDevice, c->stream));
    k_invert_orig<<<lgpu_blocks(n), LGPU_BLOCK, 0, c->stream>>>(c->orig[0], n, inv);
    k_fill_int<<<lgpu_blocks(n), LGPU_BLOCK, 0, c->stream>>>(keep, n, 1);
    int m = n_dead > n_moved ? n_dead : n_moved;
    if (m) k_apply_moves<<<lgpu_blocks(m), LGPU_BLOCK, 0, c->stream>>>(inv, d_dead, n_dead, d_moved, n_moved, keep, c->orig[0]);
    int st = lgpu_launch_scan_cells(c, keep, dst, n, false, false);
    if (st) return st;
    k_compact<<<lgpu_blocks(n), LGPU_BLOCK, 0, c->stream>>>(v, keep, dst, n);
    c->launches += 4;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    // the compacted arrays were written to buffer set 1: make it the step-boundary storage
    std::swap(c->pos[0], c->pos[1]); std::swap(c->vel[0], c->vel[1]);
    std::swap(c->flags[0], c->flags[1]); std::swap(c->orig[0], c->orig[1]);
    c->n_owned = c->n = c->n_in = n - n_dead;
    c->grid_valid = false;
    if (removed) *removed = n_dead;
    return LGPU_OK;
}

// ---------------- player AABB scan ----------------
// First k particles in reference-slot order with |p - center| <= half on every axis
// (particle_collide_with_player, src/BulletPhysics.cpp:596-600).
__global__ void k_aabb_flags(View v, int n, F3 center, F3 half, int* __restrict__ flag_by_slot) {
    int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n) return;
    float4 p = v.pos_in[u];
    bool in = fabsf(__fsub_rn(p.x, center.x)) <= half.x && fabsf(__fsub_rn(p.y, center.y)) <= half.y && fabsf(__fsub_rn(p.z, center.z)) <= half.z;
    flag_by_slot[v.orig_in[u]] = in ? 1 : 0;
}
__global__ void k_aabb_gather(View v, int n, const int* __restrict__ flag_by_slot, const int* __restrict__ rank_by_slot, const int* __restrict__ inv,
                              int k, float* __restrict__ out) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n || !flag_by_slot[s]) return;
    int r = rank_by_slot[s];
    if (r >= k) return;
    float4 p = v.pos_in[inv[s]];
    out[3 * r] = p.x; out[3 * r + 1] = p.y; out[3 * r + 2] = p.z;
}

extern "C" int lgpu_aabb_first_k(lgpu_ctx* c, const float center[3], const float half[3], int k, float* out_pos, int* out_n) {
    if (!c || !center || !half || k < 0 || !out_n || (k > 0 && !out_pos)) return LGPU_ERR_ARG;
    *out_n = 0;
    const int n = c->n_owned;
    if (n == 0 || k == 0) return LGPU_OK;
    if (k > n) k = n;  // (the output staging area holds 8 floats per particle of capacity)
    CUDA_TRY(cudaSetDevice(c->device));
    View v = lgpu_make_view(c);
    int* flag = c->key_in;   // scratch, dead between steps
    int* rank = c->rank_in;
    int* inv = c->tmp_id;
    F3 ce; ce.x = center[0]; ce.y = center[1]; ce.z = center[2];
    F3 ha; ha.x = half[0]; ha.y = half[1]; ha.z = half[2];
    k_invert_orig<<<lgpu_blocks(n), LGPU_BLOCK, 0, c->stream>>>(c->orig[0], n, inv);
    k_aabb_flags<<<lgpu_blocks(n), LGPU_BLOCK, 0, c->stream>>>(v, n, ce, ha, flag);
    int st = lgpu_launch_scan_cells(c, flag, rank, n, false, false);
    if (st) return st;
    float* d_out = c->d_stage;
    k_aabb_gather<<<lgpu_blocks(n), LGPU_BLOCK, 0, c->stream>>>(v, n, flag, rank, inv, k, d_out);
    c->launches += 3;
    CUDA_TRY(cudaGetLastError());
    int total = 0;
    CUDA_TRY(cudaMemcpyAsync(&total, rank + n, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    int m = total < k ? total : k;
    if (m > 0) CUDA_TRY(cudaMemcpy(out_pos, d_out, sizeof(float) * 3 * m, cudaMemcpyDeviceToHost));
    *out_n = m;
    return LGPU_OK;
}
