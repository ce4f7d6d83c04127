// Issue-rate microbenchmark: legacy mma.sync.m16n8k8 TF32 vs packed FFMA2 on B200 (sm_100a).
// Informs the round-2 decision on a 3xTF32 tensor-core variant of the MLP kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate tools/experiments/mma_tf32_rate.cu && ./mma_rate
#include <cstdio>
#include <cuda_runtime.h>

__global__ void mma_loop(float* out, int iters) {
    float c[8][4];
    unsigned a[4] = {0x3f800000u, 0x3f800000u, 0x3f800000u, 0x3f800000u};
    unsigned b[2] = {0x3f800000u, 0x3f800000u};
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void ffma2_loop(float* out, int iters) {
    unsigned long long acc[16], x = 0x3f8000003f800000ull, y = 0x3f0000003f000000ull;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0ull;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[i]) : "l"(x), "l"(y));
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)(s & 0xffff);
}

int main() {
    float* out;
    cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 20000;
    for (int warps = 4; warps <= 16; warps *= 2) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(a);
            mma_loop<<<148, warps * 32>>>(out, iters);
            cudaEventRecord(b); cudaEventSynchronize(b);
        }
        float ms; cudaEventElapsedTime(&ms, a, b);
        double flops = 2.0 * 16 * 8 * 8 * 8.0 * iters * warps * 148;
        printf("mma.sync m16n8k8 tf32: %2d warps/SM  %.3f ms  %.1f TFLOP/s dense\n", warps, ms, flops / (ms * 1e-3) / 1e12);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(a);
            ffma2_loop<<<148, warps * 32>>>(out, iters);
            cudaEventRecord(b); cudaEventSynchronize(b);
        }
        cudaEventElapsedTime(&ms, a, b);
        flops = 2.0 * 2 * 16.0 * iters * warps * 32 * 148;
        printf("fma.rn.f32x2        : %2d warps/SM  %.3f ms  %.1f TFLOP/s fp32\n", warps, ms, flops / (ms * 1e-3) / 1e12);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
