// Issue rate of the packed fp32 instructions of sm_100 (FFMA2 / FADD2, PTX fma.rn.f32x2 / add.rn.f32x2) against
// the scalar ones: 8 independent chains per thread, 1024 threads per SM x all SMs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/ffma2 scripts/ubench/ffma2.cu && scripts/ubench/ffma2
#include <cuda_runtime.h>
#include <stdio.h>
#define ITER 4096
__global__ void k_scalar(float *out, float a, float b) {
    float v[8];
    for (int u = 0; u < 8; ++u) v[u] = threadIdx.x + u;
    for (int it = 0; it < ITER; ++it)
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = fmaf(v[u], a, b);
    float s = 0;
    for (int u = 0; u < 8; ++u) s += v[u];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed(float *out, float a, float b) {
    unsigned long long v[8], A, B;
    float2 fa = make_float2(a, a), fb = make_float2(b, b);
    A = *reinterpret_cast<unsigned long long *>(&fa);
    B = *reinterpret_cast<unsigned long long *>(&fb);
    for (int u = 0; u < 8; ++u) { float2 t = make_float2(threadIdx.x + u, threadIdx.x - u); v[u] = *reinterpret_cast<unsigned long long *>(&t); }
    for (int it = 0; it < ITER; ++it)
#pragma unroll
        for (int u = 0; u < 8; ++u) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[u]) : "l"(A), "l"(B));
    float s = 0;
    for (int u = 0; u < 8; ++u) { float2 t = *reinterpret_cast<float2 *>(&v[u]); s += t.x + t.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mixed(float *out, float a, float b) {   // 1 FFMA2 + 1 SHF + 1 FMNMX per step: do they share issue slots?
    unsigned long long v[4], A, B;
    float2 fa = make_float2(a, a), fb = make_float2(b, b);
    A = *reinterpret_cast<unsigned long long *>(&fa);
    B = *reinterpret_cast<unsigned long long *>(&fb);
    unsigned m[4] = {0, 0, 0, 0};
    for (int u = 0; u < 4; ++u) { float2 t = make_float2(threadIdx.x + u, threadIdx.x - u); v[u] = *reinterpret_cast<unsigned long long *>(&t); }
    for (int it = 0; it < ITER; ++it)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[u]) : "l"(A), "l"(B));
            m[u] = __funnelshift_l((unsigned)(v[u] >> 32), m[u], 1);
            m[u] = __funnelshift_l((unsigned)v[u], m[u], 1);
        }
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)(m[0] ^ m[1] ^ m[2] ^ m[3]);
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *out; cudaMalloc(&out, sizeof(float) * sms * 2 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int which = 0; which < 3; ++which) {
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            if (which == 0) k_scalar<<<sms * 2, 1024>>>(out, 1.0001f, 0.5f);
            else if (which == 1) k_packed<<<sms * 2, 1024>>>(out, 1.0001f, 0.5f);
            else k_mixed<<<sms * 2, 1024>>>(out, 1.0001f, 0.5f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        const double instr = (which == 2 ? 4.0 * 3 : 8.0) * ITER * sms * 2 * 1024 / 32;  // warp instructions
        printf("%s: %.3f ms, %.2f warp-instr/clk/SM (at 1.9 GHz)%s\n", which == 0 ? "FFMA  " : which == 1 ? "FFMA2 " : "FFMA2+2SHF", best,
               instr / (best * 1e-3) / 1.9e9 / sms, which == 1 ? "  (each = 2 FMAs per lane)" : "");
    }
    return 0;
}
