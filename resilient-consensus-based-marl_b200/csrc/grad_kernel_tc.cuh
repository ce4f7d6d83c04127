// grad_kernel_tc.cuh -- EXPERIMENTAL hybrid of tensor cores and CUDA cores for the scalar-output nets (critic / team
// reward) at n_agents = 5; built only with -DRCMARL_GRAD_TC=1 (`make variant_v6`).  Written after the round-1 GPU budget
// was spent: compiles for sm_100a, NOT yet run on a GPU.  Building blocks (3xTF32 tcgen05.mma with the A operand written
// to TMEM by the row-owning threads, K-major SWIZZLE_NONE B operand, tcgen05.ld read-back) are the ones verified in
// tools/experiments/umma_tf32_probe.cu (2e-6 of fp64, 45 cycles per M128 x N32 x K8 instruction).
//
// Division of labour per 128-row tile (one "group" = 4 warps = 128 threads, thread = buffer row = TMEM lane):
//   tensor cores   z1 = [x | 1] . [W1; b1]      (K = 16, 2 k-steps x 3 split products)
//                  z2 = [h1 | 1] . [W2; b2]     (K = 24, 3 x 3)
//                  u  = delta2 . W2^T           (K = 24, 3 x 3)            24 MMAs ~ 1 080 cycles, asynchronous
//   CUDA cores     hi / lo split of the A operands, LeakyReLU, output layer (20 FMA), delta2, delta1 = u * lrelu'(h1),
//                  and phase 2 of grad_kernel.cuh (weight gradients as 8x8 register tiles over the warp's 32 tile rows).
// Phase 1 thus needs no weight traffic through the LSU and one row per lane (few registers), and the two groups of a
// CTA overlap: while one waits for its MMAs the other runs its CUDA-core part.
// TMEM columns of group g (base 256 g):  A1 hi/lo 0..31 | A2 hi/lo 32..79 | A3 hi/lo 80..127 | D1 128 | D2 160 | D3 192.
// Every wait on an MMA barrier has a clock-based bail-out that traps: a wrong descriptor must end in an error, not a hang.
#pragma once
#include "grad_kernel.cuh"
#include "tmem_ops.cuh"

namespace rcmarl {

constexpr int TC_WARPS = 8;          // two groups of four warps
constexpr int TC_ROWS = 32;          // tile rows per warp and chunk (one per lane)
constexpr int TC_N = 32;             // MMA N (20 hidden units + zero padding)

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): K-major, SWIZZLE_NONE, version 1 (Blackwell)
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A = B = tf32, both K-major
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// canonical K-major layout of the B operand: element (n, k) at (k / 4) * (32 rows * 4 floats) + n * 4 + k % 4
__host__ __device__ constexpr int tc_canon(int n, int k) { return (k >> 2) * (TC_N * 4) + n * 4 + (k & 3); }

// D[tmem_d] (+)= A[tmem_a] . B[desc]   (A from TMEM, 128 lanes x 8 tf32 columns)
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_or_trap(uint64_t* bar, uint32_t parity) {
    const long long t0 = clock64();
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!done && clock64() - t0 > 4000000000LL) __trap();     // ~2 s: the MMA never completed
    }
}
__device__ __forceinline__ void group_barrier(int group) {       // named barrier 1 + group over the group's 128 threads
    asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory");
}

// split the K values of this thread's row into tf32 hi / lo parts and store them to TMEM columns [col_hi, col_hi + K) and
// [col_lo, col_lo + K) of the thread's lane
template <int K>
__device__ __forceinline__ void tc_store_operand(uint32_t tlane, int col_hi, int col_lo, const float (&v)[K]) {
    static_assert(K % 8 == 0, "operand width");
    uint32_t h[K], l[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float hv = to_tf32(v[k]);
        h[k] = __float_as_uint(hv);
        l[k] = __float_as_uint(v[k] - hv);
    }
    tmem_store<K>(tlane + col_hi, h);
    tmem_store<K>(tlane + col_lo, l);
}

// issue the 3 x (K / 8) split products of one GEMM and commit them to the group's barrier (one thread)
template <int K>
__device__ __forceinline__ void tc_issue(uint32_t tmem_d, uint32_t tmem_a_hi, uint32_t tmem_a_lo, const float* b_hi,
                                         const float* b_lo, uint32_t idesc, uint64_t* bar) {
#pragma unroll
    for (int ks = 0; ks < K / 8; ++ks) {
        const uint32_t boff = ks * 2 * (TC_N * 16);                       // two 16-byte K chunks per k-step
        const uint64_t dbh = umma_smem_desc(smem_u32(b_hi) + boff, TC_N * 16, 128);
        const uint64_t dbl = umma_smem_desc(smem_u32(b_lo) + boff, TC_N * 16, 128);
        umma_ts(tmem_d, tmem_a_hi + ks * 8, dbh, idesc, ks > 0 ? 1u : 0u);
        umma_ts(tmem_d, tmem_a_lo + ks * 8, dbh, idesc, 1u);
        umma_ts(tmem_d, tmem_a_hi + ks * 8, dbl, idesc, 1u);
    }
    umma_commit(bar);
}

constexpr int tc_b_floats(int k) { return 2 * TC_N * k; }               // hi + lo copy of a 32 x k operand
__host__ __device__ constexpr int round32(int n) { return (n + 31) & ~31; }   // 128-byte alignment of the B operands

template <int DIN>
constexpr int grad_tc_smem_floats() {
    using L = TileLayout<DIN, 1>;
    constexpr int na = 5;
    return round32(param_count(DIN, 1)) + 32 /* alignment slack */ + tc_b_floats(L::LA1) + 2 * tc_b_floats(24) +
           TC_WARPS * TC_ROWS * L::RS + TC_WARPS * TC_ROWS * 3 * na + 2 * TC_WARPS + 2 * 2 + 8 + 16;
}

template <int NA, int DIN>
__device__ __forceinline__ void grad_body_tc(const GradParams& P, const rcmarl_grad_job& job, float* smem, int y, int gy) {
    static_assert(NA == 5, "the tensor-core variant is instantiated for n_agents = 5 only (TMEM column budget)");
    using L = TileLayout<DIN, 1>;
    constexpr int NP = param_count(DIN, 1);
    constexpr int K1 = L::LA1;                     // 16
    static_assert(K1 == 16, "operand width of layer 1");
    rcmarl_rows Rw = P.rows;
    if (job.time_idx) Rw.time_idx = job.time_idx;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = warp >> 2, gwarp = warp & 3;

    // ---- shared memory carve-up
    float* sw = smem;                                           // packed network (output layer is read from here)
    // B operands, canonical K-major, hi / lo copies, aligned to 128 bytes whatever the base of the dynamic segment is
    float* b1h = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem + NP) + 127) & ~(uintptr_t)127);
    float* b1l = b1h + TC_N * K1;
    float* b2h = b1l + TC_N * K1;
    float* b2l = b2h + TC_N * 24;
    float* b3h = b2l + TC_N * 24;
    float* b3l = b3h + TC_N * 24;
    float* tiles = b3l + TC_N * 24;
    float* wt = tiles + warp * (TC_ROWS * L::RS);
    constexpr int SROW = 3 * NA;
    float* stage = tiles + TC_WARPS * (TC_ROWS * L::RS) + warp * (TC_ROWS * SROW);
    uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + TC_WARPS * (TC_ROWS * L::RS) + TC_WARPS * (TC_ROWS * SROW));
    uint64_t* bar = bars + warp;                                // input staging (TMA) barrier of this warp
    uint64_t* mma_bar = bars + TC_WARPS + group;                // MMA completion barrier of this group
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + TC_WARPS + 2);

    if (lane == 0) mbar_init(bar, 1);
    if (threadIdx.x == 0) { mbar_init(bars + TC_WARPS, 1); mbar_init(bars + TC_WARPS + 1, 1); }
    if (lane == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (warp == 0) tmem_alloc_all(tslot);
    pdl_wait();
    stage_weights(sw, job.w, NP);
    __syncthreads();
    // B operands from the staged weights: B1[n][k] = [W1; b1][k][n], B2[n][k] = [W2; b2][k][n], B3[n][k] = W2[n][k]
    for (int i = threadIdx.x; i < TC_N * K1; i += blockDim.x) {
        const int n = i / K1, k = i % K1;
        const float v = n < HID ? (k < DIN ? sw[k * HID + n] : (k == DIN ? sw[off_b1(DIN) + n] : 0.f)) : 0.f;
        const float h = to_tf32(v);
        b1h[tc_canon(n, k)] = h;
        b1l[tc_canon(n, k)] = v - h;
    }
    for (int i = threadIdx.x; i < TC_N * 24; i += blockDim.x) {
        const int n = i / 24, k = i % 24;
        const float v2 = n < HID ? (k < HID ? sw[off_W2(DIN) + k * HID + n] : (k == HID ? sw[off_b2(DIN) + n] : 0.f)) : 0.f;
        const float v3 = (n < HID && k < HID) ? sw[off_W2(DIN) + n * HID + k] : 0.f;
        const float h2v = to_tf32(v2), h3v = to_tf32(v3);
        b2h[tc_canon(n, k)] = h2v;
        b2l[tc_canon(n, k)] = v2 - h2v;
        b3h[tc_canon(n, k)] = h3v;
        b3l[tc_canon(n, k)] = v3 - h3v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic smem writes -> async-proxy (MMA) reads
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    const uint32_t tmem = *tslot + (uint32_t)(group * 256);             // this group's column window, lane 0
    const uint32_t tlane = tmem + ((uint32_t)(gwarp * 32) << 16);       // this thread's lane in it
    constexpr int CA1H = 0, CA1L = 16, CA2H = 32, CA2L = 56, CA3H = 80, CA3L = 104, CD1 = 128, CD2 = 160, CD3 = 192;
    const uint32_t idesc = umma_idesc_tf32(128, TC_N);
    const bool issuer = (gwarp == 0) && (lane == 0);
    uint32_t mph = 0;                                                    // parity of the group's MMA barrier

    // constant columns of this lane's tile row
    {
        float* rowp = wt + lane * L::RS;
#pragma unroll
        for (int q = 0; q < L::LA1 / 4; ++q)
            if (4 * q >= DIN) st4(rowp + L::OA1 + 4 * q, 4 * q == DIN ? 1.f : 0.f, 0.f, 0.f, 0.f);
        st4(rowp + L::OA2 + 20, 1.f, 0.f, 0.f, 0.f);
        st4(rowp + L::OD1 + 20, 0.f, 0.f, 0.f, 0.f);
        st4(rowp + L::OD2 + 20, 0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();

    const bool from_ns = (DIN == 2 * NA) && job.kind == RCMARL_IN_NS;
    const int rowf = from_ns ? 2 * NA : 3 * NA;
    const float* in_base = from_ns ? Rw.ns : Rw.sa;
    const bool gather_ok = (Rw.time_idx == nullptr) || (Rw.n_envs % 128 == 0);
    // the 32 rows of this warp inside tile T: 128 T + 32 gwarp + lane
    auto stage_src = [&](int64_t T, const float*& src) -> bool {
        const int64_t m0 = T * 128 + gwarp * TC_ROWS;
        if (!gather_ok || m0 + TC_ROWS > Rw.n_rows) return false;
        src = in_base + row_of(Rw, m0) * rowf;
        return (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    };
    uint32_t phase = 0;
    bool staged = false;
    const int64_t ntiles = (Rw.n_rows + 127) / 128;
    const int64_t tstep = (int64_t)gy * 2;
    const int64_t tfirst = (int64_t)group * gy + y;
    {
        const float* src = nullptr;
        if (tfirst < ntiles) staged = stage_src(tfirst, src);
        if (staged && lane == 0) bulk_load(stage, src, (uint32_t)(TC_ROWS * rowf * sizeof(float)), bar);
    }

    // phase-2 assignment of this lane (as in grad_body)
    const bool busy = lane < L::NG * L::NT;
    const int tile = busy ? lane % L::NT : 0;
    const int grp = busy ? lane / L::NT : 0;
    int aoff, doff;
    L::tile_offsets(tile, aoff, doff);
    f2 acc[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) acc[e] = pack2(0.f, 0.f);
    float g3[HID + 1];
#pragma unroll
    for (int j = 0; j <= HID; ++j) g3[j] = 0.f;
    float loss = 0.f;

    for (int64_t T = tfirst; T < ntiles; T += tstep) {
        const int64_t m = T * 128 + gwarp * TC_ROWS + lane;
        const bool live = m < Rw.n_rows;
        const int64_t row = row_of(Rw, live ? m : 0);                   // dead rows read row 0 and contribute zeros
        float* rowp = wt + lane * L::RS;
        float h1[HID];
        // ---------------- layer 1 on the tensor cores ----------------
        {
            float x[K1];
            if (staged) {
                mbar_wait(bar, phase);
                phase ^= 1u;
                const int skip = (DIN == 2 * NA && !from_ns) ? 1 : 0;
                const float* sp = stage + lane * rowf;
#pragma unroll
                for (int k = 0; k < DIN; ++k) x[k] = sp[k + skip * (k >> 1)];
            } else {
                float xr[DIN];
                load_x<NA, DIN>(Rw, job.kind, row, xr);
#pragma unroll
                for (int k = 0; k < DIN; ++k) x[k] = xr[k];
            }
#pragma unroll
            for (int k = DIN; k < K1; ++k) x[k] = (k == DIN) ? 1.f : 0.f;
            __syncwarp();
            {
                const int64_t T2 = T + tstep;
                const float* src = nullptr;
                staged = (T2 < ntiles) && stage_src(T2, src);
                if (staged && lane == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    bulk_load(stage, src, (uint32_t)(TC_ROWS * rowf * sizeof(float)), bar);
                }
            }
#pragma unroll
            for (int q = 0; q < L::LA1 / 4; ++q)
                if (4 * q < DIN) st4(rowp + L::OA1 + 4 * q, x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
            tc_store_operand<K1>(tlane, CA1H, CA1L, x);
            tmem_wait_st();
            tmem_fence_before_sync();
            group_barrier(group);
            if (issuer) {
                tmem_fence_after_sync();
                tc_issue<K1>(tmem + CD1, tmem + CA1H, tmem + CA1L, b1h, b1l, idesc, mma_bar);
            }
            mbar_wait_or_trap(mma_bar, mph);
            mph ^= 1u;
            tmem_fence_after_sync();
            uint32_t z[24];
            tmem_load<24>(tlane + CD1, z);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < HID; ++j) {
                const float zz = __uint_as_float(z[j]);
                h1[j] = fmaxf(zz, SLOPE * zz);
            }
        }
        // ---------------- layer 2 ----------------
        float d2[HID];
        {
            float a2[24];
#pragma unroll
            for (int j = 0; j < HID; ++j) a2[j] = h1[j];
            a2[20] = 1.f; a2[21] = 0.f; a2[22] = 0.f; a2[23] = 0.f;
#pragma unroll
            for (int q = 0; q < 5; ++q) st4(rowp + L::OA2 + 4 * q, h1[4 * q], h1[4 * q + 1], h1[4 * q + 2], h1[4 * q + 3]);
            tc_store_operand<24>(tlane, CA2H, CA2L, a2);
            tmem_wait_st();
            tmem_fence_before_sync();
            group_barrier(group);
            if (issuer) {
                tmem_fence_after_sync();
                tc_issue<24>(tmem + CD2, tmem + CA2H, tmem + CA2L, b2h, b2l, idesc, mma_bar);
            }
            mbar_wait_or_trap(mma_bar, mph);
            mph ^= 1u;
            tmem_fence_after_sync();
            uint32_t z[24];
            tmem_load<24>(tlane + CD2, z);
            tmem_wait_ld();
            float h2[HID];
#pragma unroll
            for (int j = 0; j < HID; ++j) {
                const float zz = __uint_as_float(z[j]);
                h2[j] = fmaxf(zz, SLOPE * zz);
            }
            // output layer, loss and delta2 on the CUDA cores (Keras MSE: dLoss/dout = 2 (out - y) / B, 2/B by the caller)
            const float tgt = live ? __ldg(job.target + row * job.target_stride) : 0.f;
            const SmemW W{sw};
            const float e = live ? head1_w<DIN>(W, h2) - tgt : 0.f;
            loss = fmaf(e, e, loss);
#pragma unroll
            for (int j = 0; j < HID; ++j) {
                g3[j] = fmaf(h2[j], e, g3[j]);
                d2[j] = W.s(off_W3(DIN) + j) * e * lrelu_grad_from_out(h2[j]);
            }
            g3[HID] += e;
        }
        // ---------------- backward-data: u = delta2 . W2^T ----------------
        {
            float a3[24];
#pragma unroll
            for (int j = 0; j < HID; ++j) a3[j] = d2[j];
            a3[20] = 0.f; a3[21] = 0.f; a3[22] = 0.f; a3[23] = 0.f;
#pragma unroll
            for (int q = 0; q < 5; ++q) st4(rowp + L::OD2 + 4 * q, d2[4 * q], d2[4 * q + 1], d2[4 * q + 2], d2[4 * q + 3]);
            tc_store_operand<24>(tlane, CA3H, CA3L, a3);
            tmem_wait_st();
            tmem_fence_before_sync();
            group_barrier(group);
            if (issuer) {
                tmem_fence_after_sync();
                tc_issue<24>(tmem + CD3, tmem + CA3H, tmem + CA3L, b3h, b3l, idesc, mma_bar);
            }
            mbar_wait_or_trap(mma_bar, mph);
            mph ^= 1u;
            tmem_fence_after_sync();
            uint32_t u[24];
            tmem_load<24>(tlane + CD3, u);
            tmem_wait_ld();
            float d1[HID];
#pragma unroll
            for (int i = 0; i < HID; ++i) d1[i] = __uint_as_float(u[i]) * lrelu_grad_from_out(h1[i]);
#pragma unroll
            for (int q = 0; q < 5; ++q) st4(rowp + L::OD1 + 4 * q, d1[4 * q], d1[4 * q + 1], d1[4 * q + 2], d1[4 * q + 3]);
            tmem_fence_before_sync();       // the next tile's MMAs overwrite D1..D3 only after the group barrier of layer 1
        }
        __syncwarp();
        // ---------------- phase 2: 8x8 register tile per lane over the warp's 32 tile rows ----------------
#pragma unroll 4
        for (int it = 0; it < TC_ROWS / L::NG; ++it) {
            const float* rp = wt + (it * L::NG + grp) * L::RS;
            const float4 a0 = *reinterpret_cast<const float4*>(rp + aoff);
            const float4 a1 = *reinterpret_cast<const float4*>(rp + aoff + 4);
            const float4 d0 = *reinterpret_cast<const float4*>(rp + doff);
            const float4 dd1 = *reinterpret_cast<const float4*>(rp + doff + 4);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const f2 d[4] = {pack2(d0.x, d0.y), pack2(d0.z, d0.w), pack2(dd1.x, dd1.y), pack2(dd1.z, dd1.w)};
#pragma unroll
            for (int ii = 0; ii < 8; ++ii) {
                const f2 aa = pack2(a[ii], a[ii]);
#pragma unroll
                for (int jp = 0; jp < 4; ++jp) acc[ii * 4 + jp] = fma2(aa, d[jp], acc[ii * 4 + jp]);
            }
        }
        __syncwarp();
    }

    // ---------------- CTA reduction (fixed order => bitwise reproducible), as in grad_body ----------------
#if RCMARL_PDL_REDUCE
    pdl_launch_dependents();
#endif
    tmem_fence_before_sync();
    __syncthreads();
    float* red = tiles;                               // [TC_WARPS][32][64]
    {
        float4* dst = reinterpret_cast<float4*>(red + (warp * 32 + lane) * 64);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            float4 v;
            unpack2(acc[2 * q], v.x, v.y);
            unpack2(acc[2 * q + 1], v.z, v.w);
            dst[q] = v;
        }
    }
    float* red3 = red + TC_WARPS * 32 * 64;
    loss = warp_sum(loss);
    if (lane == 0) red3[warp * (HID + 2) + HID + 1] = loss;
#pragma unroll
    for (int j = 0; j <= HID; ++j) {
        const float s = warp_sum(g3[j]);
        if (lane == 0) red3[warp * (HID + 2) + j] = s;
    }
    __syncthreads();
    float* out = P.partial + (int64_t)blockIdx.x * P.stride;
    for (int q = threadIdx.x; q < L::NT * 64; q += blockDim.x) {
        const int t = q >> 6, e = q & 63;
        const int idx = L::tile_param(t, e >> 3, e & 7);
        if (idx >= 0) {
            float s = 0.f;
            for (int w = 0; w < TC_WARPS; ++w)
#pragma unroll
                for (int g = 0; g < L::NG; ++g) s += red[(w * 32 + g * L::NT + t) * 64 + e];
            out[idx] = s;
        }
    }
    if (threadIdx.x <= HID) {
        float s = 0.f;
        for (int w = 0; w < TC_WARPS; ++w) s += red3[w * (HID + 2) + threadIdx.x];
        out[(threadIdx.x < HID ? off_W3(DIN) : off_b3(DIN, 1) - HID) + threadIdx.x] = s;
    }
    if (threadIdx.x == 32) {
        float s = 0.f;
        for (int w = 0; w < TC_WARPS; ++w) s += red3[w * (HID + 2) + HID + 1];
        out[NP] = s;
    }
    if (warp == 0) tmem_dealloc_all(*tslot);
}

// scalar-output nets at n_agents = 5 only (mean-squared-error jobs); everything else stays on grad_kernel
__global__ void __launch_bounds__(32 * TC_WARPS, 1) grad_kernel_tc(const __grid_constant__ GradParams P) {
    extern __shared__ __align__(16) float smem[];
    int j = 0;
    while (j + 1 < P.n_jobs && (int)blockIdx.x >= P.cta_first[j + 1]) ++j;
    const rcmarl_grad_job& job = P.jobs[j];
    const int y = (int)blockIdx.x - P.cta_first[j], gy = P.cta_first[j + 1] - P.cta_first[j];
    if (job.kind == RCMARL_IN_SA) grad_body_tc<5, 15>(P, job, smem, y, gy);
    else grad_body_tc<5, 10>(P, job, smem, y, gy);
}

}  // namespace rcmarl
