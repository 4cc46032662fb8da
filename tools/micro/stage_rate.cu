// Micro-benchmark: how fast can ONE SM fill shared memory from L2-resident global memory with many small 1-D ranges?
//   mode 0: cp.async.bulk (TMA), one copy per lane per round (divergent lanes, as the brick producer issues them)
//   mode 1: cp.async 16 B per lane (LDGSTS), `warps` producer warps, each range copied by a whole warp
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stage_rate stage_rate.cu && ./stage_rate
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
    uint32_t ok;
    do { asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(s32(b)), "r"(par), "r"(20000u) : "memory"); } while (!ok);
}
__device__ __forceinline__ void bulk(void* d, const void* s, uint32_t bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(d)), "l"(s), "r"(bytes), "r"(s32(b)) : "memory");
}
__device__ __forceinline__ void cpa16(void* d, const void* s) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(d)), "l"(s) : "memory"); }
__device__ __forceinline__ void cpa_arrive(uint64_t* b) { asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(s32(b)) : "memory"); }

// every block: `rounds` bricks, each = ncopy ranges of `bytes` bytes (ranges 4 KB apart in global memory)
__global__ void __launch_bounds__(1024, 1) k(const float4* src, size_t src_elems, int mode, int warps, int ncopy, int bytes, int rounds, long long* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* bar = (uint64_t*)sm;
    float4* dst = (float4*)(sm + 128);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(bar, mode != 1 ? 1 : 32 * warps); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (warp >= warps) return;
    const int per = bytes / 16;
    long long t_issue = 0, t_total = 0;
    const size_t base0 = ((size_t)blockIdx.x * 7919u * 256u) % (src_elems - (size_t)ncopy * 256 - 4096);
    for (int r = 0; r < rounds; r++) {
        const size_t base = (base0 + (size_t)r * 104729u * 64u) % (src_elems - (size_t)ncopy * 256 - 4096);
        long long t0 = clock64();
        if (mode == 0) {
            if (lane == 0) mbar_expect(bar, (uint32_t)(ncopy * bytes));
            __syncwarp();
            for (int c = lane; c < ncopy; c += 32) bulk(dst + (size_t)c * per, src + base + (size_t)c * 256, bytes, bar);
        } else if (mode == 2) {   // bulk copies split among `warps` warps, one lane each per round
            if (warp == 0 && lane == 0) mbar_expect(bar, (uint32_t)(ncopy * bytes));
            asm volatile("bar.sync 1, %0;" ::"r"(32 * warps));
            for (int c = warp * 32 + lane; c < ncopy; c += 32 * warps) bulk(dst + (size_t)c * per, src + base + (size_t)c * 256, bytes, bar);
        } else if (mode == 3) {   // bulk copies issued by lane 0 in a uniform loop
            if (lane == 0) {
                mbar_expect(bar, (uint32_t)(ncopy * bytes));
                for (int c = 0; c < ncopy; c++) bulk(dst + (size_t)c * per, src + base + (size_t)c * 256, bytes, bar);
            }
        } else if (mode == 4) {   // one lane of each of `warps` warps, uniform loops
            if (warp == 0 && lane == 0) mbar_expect(bar, (uint32_t)(ncopy * bytes));
            asm volatile("bar.sync 1, %0;" ::"r"(32 * warps));
            if (lane == 0) for (int c = warp; c < ncopy; c += warps) bulk(dst + (size_t)c * per, src + base + (size_t)c * 256, bytes, bar);
        } else {
            for (int c = warp; c < ncopy; c += warps)
                for (int o = lane; o < per; o += 32) cpa16(dst + (size_t)c * per + o, src + base + (size_t)c * 256 + o);
            cpa_arrive(bar);
        }
        long long t1 = clock64();
        mbar_wait(bar, r & 1);
        long long t2 = clock64();
        t_issue += t1 - t0; t_total += t2 - t0;
    }
    if (threadIdx.x == 0) { out[2 * blockIdx.x] = t_issue / rounds; out[2 * blockIdx.x + 1] = t_total / rounds; }
}
int main() {
    const size_t elems = (size_t)(getenv("BIG") ? 96 : 4) << 20;  // 64 MB of float4: L2 resident after the first touch (BIG: 1.5 GB, DRAM)
    float4* src; cudaMalloc(&src, elems * 16); cudaMemset(src, 0, elems * 16);
    long long* out; cudaMalloc(&out, 148 * 2 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    long long h[296];
    struct { int mode, warps, ncopy, bytes; } cases[] = {
        {0, 1, 36, 592}, {0, 1, 72, 592}, {0, 1, 16, 2048}, {0, 1, 9, 4096}, {0, 1, 1, 32768}, {0, 1, 32, 1024}, {0, 1, 1, 592}, {0,1,4,592},
        {4, 4, 36, 592}, {4, 8, 36, 592}, {4, 4, 18, 1184}, {4, 4, 9, 2368}, {4, 1, 1, 21312}, {4, 4, 4, 8192},
    };
    for (auto& c : cases) {
        for (int grid : {1, 148}) {
            k<<<grid, 1024, 200 * 1024>>>(src, elems, c.mode, c.warps, c.ncopy, c.bytes, 50, out);  // warm
            k<<<grid, 1024, 200 * 1024>>>(src, elems, c.mode, c.warps, c.ncopy, c.bytes, 200, out);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            double bytes = (double)c.ncopy * c.bytes;
            printf("%s warps %d  %3d x %5d B (%6.0f B)  grid %3d: issue %6lld cyc, total %6lld cyc  -> %.1f B/cyc/SM  %s\n", c.mode == 1 ? "LDGSTS" : c.mode == 0 ? "BULK  " : c.mode == 2 ? "BULKmw" : c.mode == 3 ? "BULKl0" : "BULKw0", c.warps, c.ncopy, c.bytes, bytes,
                   grid, h[0], h[1], bytes / (double)h[1], e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    }
    return 0;
}
