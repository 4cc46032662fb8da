// lgpu_neighbors.cu — builds the per-particle neighbour table once per substep.
// Replaces the list construction of find_neighbors_uniform_grid (src/neighbors/Neighbors.cpp:386-448)
// and find_neighbors_uniform_grid_v1 (:306-361).
#include "lgpu_neighbors.cuh"

template <bool SAND>
__global__ void __launch_bounds__(LGPU_BLOCK) k_build_table(View v) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.n) return;
    int cnt = 0;
    int* col = v.nbr + i;
    const int M = v.M;
    const size_t stride = (size_t)v.cap;
    walk<SAND>(v, i, f3(v.x0[i]), [&](int j) {
        if (cnt < M) col[(size_t)cnt * stride] = j;
        cnt++;
    });
    v.nbr_cnt[i] = cnt;
    if (cnt > M) atomicAdd(&v.counters[1], 1ULL);
}

int lgpu_launch_build_table(lgpu_ctx* c, bool sand_order) {
    if (c->n == 0) return LGPU_OK;
    View v = lgpu_make_view(c);
    if (sand_order) k_build_table<true><<<lgpu_blocks(c->n), LGPU_BLOCK, 0, c->stream>>>(v);
    else k_build_table<false><<<lgpu_blocks(c->n), LGPU_BLOCK, 0, c->stream>>>(v);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return LGPU_OK;
}
