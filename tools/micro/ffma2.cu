// Does FFMA2 (fma.rn.f32x2) double the fp32 FMA rate per issue slot on sm_100a?  16 independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, int iters) {
    float a[16]; unsigned long long p[8];
    for (int i = 0; i < 16; i++) a[i] = threadIdx.x * 0.001f + i;
    for (int i = 0; i < 8; i++) { float2 t = make_float2(a[2 * i], a[2 * i + 1]); p[i] = *reinterpret_cast<unsigned long long*>(&t); }
    const float m = 1.0001f, c = 0.5f;
    float2 mm = make_float2(m, m), cc = make_float2(c, c);
    unsigned long long m2 = *reinterpret_cast<unsigned long long*>(&mm), c2 = *reinterpret_cast<unsigned long long*>(&cc);
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], m, c);
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(m2), "l"(c2));
        }
    }
    float s = 0;
    if (MODE == 0) for (int i = 0; i < 16; i++) s += a[i];
    else for (int i = 0; i < 8; i++) { float2 t = *reinterpret_cast<float2*>(&p[i]); s += t.x + t.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 2 * 512 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; mode++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 2, 512>>>(out, iters); else k<1><<<148 * 2, 512>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double fma = 148.0 * 2 * 512 * 16.0 * iters;
            if (rep) printf("%s: %.3f ms  %.1f TFMA/s  (%.1f TFLOP/s)\n", mode ? "FFMA2" : "FFMA ", ms, fma / ms * 1e-9, 2 * fma / ms * 1e-9);
        }
    }
    return 0;
}
