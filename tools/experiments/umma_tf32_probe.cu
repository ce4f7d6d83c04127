// tcgen05 probe (round-2 groundwork): one 128 x 32 x 24 GEMM tile D = A . B^T as three error-compensated
// TF32 products (A_hi.B_hi + A_lo.B_hi + A_hi.B_lo, "3xTF32") with tcgen05.mma.kind::tf32, operands in shared memory
// (K-major, no swizzle, canonical 8-row x 16-byte core matrices), accumulator in TMEM, read back with tcgen05.ld.
// Checks the result against fp64 and measures the issue rate of the 9-MMA group.  Every wait has a clock-based bail-out:
// a wrong descriptor must produce a wrong number or an error code, never a hung GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe tools/experiments/umma_tf32_probe.cu && timeout 30 ./umma_probe
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int M = 128, N = 32, K = 24, KSTEP = 8, NKSTEP = K / KSTEP;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): K-major, SWIZZLE_NONE, version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A = B = tf32, both K-major
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ bool mbar_wait_timeout(uint64_t* bar, uint32_t parity) {
    const long long t0 = clock64();
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!done && clock64() - t0 > 2000000000LL) return false;     // ~1 s
    }
    return true;
}

// canonical K-major layout used here: element (row r, k) lives at  kchunk * (ROWS*16 B) + r * 16 B + (k % 4) * 4 B,
// kchunk = k / 4: core matrix = 8 consecutive rows x 16 B (128 B contiguous), SBO = 128 B between 8-row groups,
// LBO = ROWS*16 B between the two 16-byte K chunks of one MMA (K = 8 tf32)
template <int ROWS>
__device__ __forceinline__ int canon(int r, int k) { return (k >> 2) * (ROWS * 4) + r * 4 + (k & 3); }

// mode 0: A from shared memory (SS);  mode 1: A from TMEM (TS): every thread stores its own row's hi / lo values with
// tcgen05.st.32x32b (lane = row, consecutive registers = consecutive columns) -- the thread-per-row layout of the MLP kernels
template <int mode>
__global__ void __launch_bounds__(128) probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
                                             int reps, int* status, long long* cycles) {
    __shared__ __align__(128) float a_hi[M * K], a_lo[M * K], b_hi[N * K], b_lo[N * K];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    const float* Ablk = A + (size_t)blockIdx.x * M * K;
    for (int k = 0; k < K; ++k) {                                       // thread = row of A
        const float v = Ablk[tid * K + k], h = to_tf32(v);
        a_hi[canon<M>(tid, k)] = h;
        a_lo[canon<M>(tid, k)] = v - h;
    }
    for (int i = tid; i < N * K; i += 128) {
        const int n = i / K, k = i % K;
        const float v = B[i], h = to_tf32(v);
        b_hi[canon<N>(n, k)] = h;
        b_lo[canon<N>(n, k)] = v - h;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {                                                     // one warp allocates 32 TMEM columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic smem writes -> async-proxy (MMA) reads
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    const uint32_t idesc = make_idesc(M, N);
    const uint32_t tmem_ahi = tmem + 32, tmem_alo = tmem + 32 + K;       // columns 32..55 and 56..79
    if (mode == 1) {
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        for (int ks = 0; ks < NKSTEP; ++ks) {
            uint32_t h[8], l[8];
            for (int u = 0; u < 8; ++u) {
                const float v = Ablk[tid * K + ks * 8 + u], hv = to_tf32(v);
                h[u] = __float_as_uint(hv);
                l[u] = __float_as_uint(v - hv);
            }
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                         ::"r"(tmem_ahi + lane_base + ks * 8), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]), "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7]) : "memory");
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                         ::"r"(tmem_alo + lane_base + ks * 8), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]), "r"(l[4]), "r"(l[5]), "r"(l[6]), "r"(l[7]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    long long t0 = 0, t1 = 0;
    bool ok = true;
    uint32_t parity = 0;
    if (tid == 0) {
        // descriptors are loop invariants: the timed loop is nothing but the 9 tcgen05.mma issues per repetition
        uint64_t dah[NKSTEP], dal[NKSTEP], dbh[NKSTEP], dbl[NKSTEP];
        for (int ks = 0; ks < NKSTEP; ++ks) {
            const uint32_t aoff = ks * 2 * (M * 16), boff = ks * 2 * (N * 16);           // two 16-byte K chunks per step
            dah[ks] = make_desc(smem_u32(a_hi) + aoff, M * 16, 128); dal[ks] = make_desc(smem_u32(a_lo) + aoff, M * 16, 128);
            dbh[ks] = make_desc(smem_u32(b_hi) + boff, N * 16, 128); dbl[ks] = make_desc(smem_u32(b_lo) + boff, N * 16, 128);
        }
        t0 = clock64();
#pragma unroll 1
        for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
            for (int ks = 0; ks < NKSTEP; ++ks) {
                const uint32_t acc0 = (ks > 0) ? 1u : 0u;                                  // first MMA of a rep overwrites D
                if (mode == 2) {
                    // three INDEPENDENT accumulators (TMEM columns 0, 32, 64), issue interleaved: is the ~47-cycle interval of
                    // modes 0/1 a dependency latency on the shared accumulator or the pipe's throughput for this shape?
#pragma unroll
                    for (int which = 0; which < 3; ++which)
#pragma unroll
                        for (int t = 0; t < 3; ++t) {
                            const uint64_t da = which == 1 ? dal[ks] : dah[ks], db = which == 2 ? dbl[ks] : dbh[ks];
                            const uint32_t acc = (ks > 0 || which > 0) ? 1u : 0u;
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                                         ::"r"(tmem + 32 * t), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
                        }
                } else if (mode == 1) {
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                                 ::"r"(tmem), "r"(tmem_ahi + ks * 8), "l"(dbh[ks]), "r"(idesc), "r"(acc0) : "memory");
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                                 ::"r"(tmem), "r"(tmem_alo + ks * 8), "l"(dbh[ks]), "r"(idesc), "r"(1u) : "memory");
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                                 ::"r"(tmem), "r"(tmem_ahi + ks * 8), "l"(dbl[ks]), "r"(idesc), "r"(1u) : "memory");
                } else {
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem), "l"(dah[ks]), "l"(dbh[ks]), "r"(idesc), "r"(acc0) : "memory");
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem), "l"(dal[ks]), "l"(dbh[ks]), "r"(idesc), "r"(1u) : "memory");
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem), "l"(dah[ks]), "l"(dbl[ks]), "r"(idesc), "r"(1u) : "memory");
                }
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    ok = mbar_wait_timeout(&bar, parity);
    if (tid == 0) t1 = clock64();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (ok) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);      // warp w owns TMEM lanes 32w .. 32w+31
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                       "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                       "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int n = 0; n < N; ++n) C[((size_t)blockIdx.x * M + tid) * N + n] = __uint_as_float(v[n]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
    if (tid == 0) {
        if (!ok) atomicExch(status, 1);
        if (blockIdx.x == 0) *cycles = t1 - t0;
    }
}

int main() {
    const int blocks_max = 148;
    float *hA = (float*)malloc(sizeof(float) * blocks_max * M * K), *hB = (float*)malloc(sizeof(float) * N * K);
    srand(1);
    for (int i = 0; i < blocks_max * M * K; ++i) hA[i] = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (int i = 0; i < N * K; ++i) hB[i] = (float)rand() / RAND_MAX * 2.f - 1.f;
    float *dA, *dB, *dC;
    int* dstatus;
    long long* dcyc;
    cudaMalloc(&dA, sizeof(float) * blocks_max * M * K);
    cudaMalloc(&dB, sizeof(float) * N * K);
    cudaMalloc(&dC, sizeof(float) * blocks_max * M * N);
    cudaMalloc(&dstatus, sizeof(int));
    cudaMalloc(&dcyc, sizeof(long long));
    cudaMemcpy(dA, hA, sizeof(float) * blocks_max * M * K, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB, sizeof(float) * N * K, cudaMemcpyHostToDevice);
    cudaMemset(dstatus, 0, sizeof(int));
    cudaMemset(dC, 0, sizeof(float) * blocks_max * M * N);
    cudaError_t e;
    int st = 0;
    float* hC = (float*)malloc(sizeof(float) * M * N);
    const int reps = 4000;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int mode = 0; mode < 3; ++mode) {
    printf("---- mode %d: %s\n", mode, mode == 2 ? "SS, three independent accumulators issued interleaved (27 MMAs per repetition)"
                                       : mode ? "A operand from TMEM (TS, tcgen05.st by the row-owning thread)" : "A operand from shared memory (SS)");
    cudaMemset(dC, 0, sizeof(float) * blocks_max * M * N);
    if (mode == 2) probe<2><<<1, 128>>>(dA, dB, dC, 1, dstatus, dcyc); else if (mode) probe<1><<<1, 128>>>(dA, dB, dC, 1, dstatus, dcyc); else probe<0><<<1, 128>>>(dA, dB, dC, 1, dstatus, dcyc);
    e = cudaDeviceSynchronize();
    cudaMemcpy(&st, dstatus, sizeof(int), cudaMemcpyDeviceToHost);
    printf("correctness launch: %s, timeout flag %d\n", cudaGetErrorString(e), st);
    if (e != cudaSuccess || st) return 1;
    cudaMemcpy(hC, dC, sizeof(float) * M * N, cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int r = 0; r < M; ++r)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)hA[r * K + k] * (double)hB[n * K + k];
            worst = fmax(worst, fabs(ref - hC[r * N + n]));
        }
    printf("3xTF32 128x32x24 tile vs fp64: max abs err %.3e (plain fp32 accumulation would give ~1e-6; 1xTF32 ~1e-3)\n", worst);
    // issue-rate: many repetitions of the 9-MMA group on every SM
    for (int blocks : {1, 148}) {
        cudaEventRecord(a);
        if (mode == 2) probe<2><<<blocks, 128>>>(dA, dB, dC, reps, dstatus, dcyc); else if (mode) probe<1><<<blocks, 128>>>(dA, dB, dC, reps, dstatus, dcyc); else probe<0><<<blocks, 128>>>(dA, dB, dC, reps, dstatus, dcyc);
        cudaEventRecord(b);
        e = cudaDeviceSynchronize();
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        cudaMemcpy(&st, dstatus, sizeof(int), cudaMemcpyDeviceToHost);
        long long cyc = 0;
        cudaMemcpy(&cyc, dcyc, sizeof(cyc), cudaMemcpyDeviceToHost);
        const double n_mma = (mode == 2 ? 27.0 : 9.0);
        const double flop = 2.0 * M * N * KSTEP * n_mma * reps * blocks;
        printf("%3d CTA(s): %s timeout %d  %.3f ms  %.1f TFLOP/s dense tf32 (= %.1f TFLOP/s useful fp32 after the 3x split), "
               "%.1f cycles per 128x32x8 MMA\n", blocks, cudaGetErrorString(e), st, ms, flop / (ms * 1e-3) / 1e12,
               flop / 3.0 / (ms * 1e-3) / 1e12, (double)cyc / (n_mma * reps));
    }
    }
    return 0;
}
