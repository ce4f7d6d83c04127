// grad_kernel_v4.cuh -- EXPERIMENTAL successor of grad_body (grad_kernel.cuh); built only with -DRCMARL_GRAD_V4=1
// (`make variant_v5`), written after the round-1 GPU budget was spent: compiles for sm_100a, NOT yet run on a GPU.
//
// Why: in grad_body the 64 accumulator registers of the 8x8 weight-gradient tile stay live across phase 1, which caps
// both phases at 4 FMA per shared-memory wavefront.  Here the accumulators are parked in Tensor
// Memory (TMEM, 256 KB per SM, reachable only through tcgen05.ld / tcgen05.st; each thread owns one TMEM lane) while
// phase 1 runs, so phase 2 can use exact-fit tiles without spilling:
//   pass A: layer-2 gradient [h1 | 1] x delta2 as four 12 x 10 tiles (+ two for the actor's output layer),
//           120 accumulators per lane, 8 row groups;  3 LDS.128 + 5 LDS.64 = 22 wavefronts per 60 FFMA2;
//   pass B: layer-1 gradient [x | 1] x delta1 as 8 x 10 tiles, 80 accumulators per lane;
//           2 LDS.128 + 5 LDS.64 = 18 wavefronts per 40 FFMA2.
// Against the 8x8 tiles (16 wavefronts per 32 FFMA2, 960 padded FMA per row) that is 800 FMA per row (team-reward net)
// and 5 / 8 of the wavefronts.  Cost: 200 registers per lane read back from TMEM per 64-row chunk (LDTM: 64 B/clk/SM).
// Phase 1 is the code of grad_body with the new tile-row layout.
#pragma once
#include "grad_kernel.cuh"
#include "tmem_ops.cuh"

namespace rcmarl {

template <int DIN, int NOUT>
struct TileLayout4 {
    static constexpr bool L3T = (NOUT > 1);
    static constexpr int LA1 = round8(DIN + 1);          // [x .., 1, 0 pad]
    static constexpr int OA1 = 0;
    static constexpr int OA2 = LA1;                      // [h1(20), 1, 0, 0, 0]
    static constexpr int OA3 = OA2 + 24;                 // [h2(20), 1, 0, 0, 0]   (actor only)
    static constexpr int OD1 = OA3 + (L3T ? 24 : 0);     // delta1 (20), no padding: 10-wide tile columns
    static constexpr int OD2 = OD1 + 20;                 // delta2 (20)
    static constexpr int OD3 = OD2 + 20;                 // actor: dLoss/dlogits (5) + 7 zeros
    static constexpr int RAW = OD3 + (L3T ? 12 : 0);
    // row stride: odd number of 16-byte units => conflict-free float4 stores within a quarter-warp; with it the 64-bit
    // delta loads of four different rows and two column offsets also fall into distinct banks
    static constexpr int RS = ((RAW / 4) % 2 == 0) ? RAW + 4 : RAW;
    static constexpr int NTA = 4 + (L3T ? 2 : 0);        // 12 x 10 tiles of pass A
    static constexpr int NGA = 32 / NTA;                 // row groups of pass A
    static constexpr int NTB = (LA1 / 8) * 2;            // 8 x 10 tiles of pass B
    static constexpr int NGB = 32 / NTB;
    static_assert(NTB <= 32, "tile count exceeds a warp");
    static constexpr int ROWS = 64;
    static constexpr int ACC_A = 120, ACC_B = 80;        // accumulators per lane (TMEM columns)

    // shared-memory offsets of the a- and delta-fragments of a pass-A / pass-B tile
    __host__ __device__ static constexpr int a_off_A(int t) { return t < 4 ? OA2 + 12 * (t / 2) : OA3 + 12 * (t - 4); }
    __host__ __device__ static constexpr int d_off_A(int t) { return t < 4 ? OD2 + 10 * (t % 2) : OD3; }
    __host__ __device__ static constexpr int a_off_B(int t) { return OA1 + 8 * (t / 2); }
    __host__ __device__ static constexpr int d_off_B(int t) { return OD1 + 10 * (t % 2); }

    // packed-parameter index of element (ii, jj) of a tile, -1 for padding
    __host__ __device__ static constexpr int param_A(int t, int ii, int jj) {
        if (t < 4) {
            const int i = 12 * (t / 2) + ii, j = 10 * (t % 2) + jj;
            return i < HID ? off_W2(DIN) + i * HID + j : (i == HID ? off_b2(DIN) + j : -1);
        }
        const int i = 12 * (t - 4) + ii, o = jj;
        if (o >= NOUT) return -1;
        return i < HID ? off_W3(DIN) + i * NOUT + o : (i == HID ? off_b3(DIN, NOUT) + o : -1);
    }
    __host__ __device__ static constexpr int param_B(int t, int ii, int jj) {
        const int i = 8 * (t / 2) + ii, j = 10 * (t % 2) + jj;
        return i < DIN ? i * HID + j : (i == DIN ? off_b1(DIN) + j : -1);
    }
};

template <int DIN, int NOUT, bool SA_NET>
constexpr int grad4_warps_for() {
    using L = TileLayout4<DIN, NOUT>;
    const int na = SA_NET ? DIN / 3 : DIN / 2;
    const int avail = GRAD_SMEM_BUDGET - 4 * (round4(param_count(DIN, NOUT)) + 32) - 128;
    const int n = avail / (4 * L::ROWS * L::RS + (grad_use_tma(na) ? 4 * L::ROWS * stage_floats_per_row(DIN, SA_NET) : 0));
    return n > 8 ? 8 : n;
}
template <int NA, int LOSS>
constexpr int grad4_warps() {
    if (LOSS == RCMARL_LOSS_CE) return grad4_warps_for<2 * NA, NACT, false>();
    const int a = grad4_warps_for<3 * NA, 1, true>(), b = grad4_warps_for<2 * NA, 1, false>();
    return a < b ? a : b;
}
template <int NA, int DIN, int NOUT, int NW>
constexpr int grad4_smem_floats() {
    using L = TileLayout4<DIN, NOUT>;
    constexpr int tiles = NW * L::ROWS * L::RS;
    constexpr int red = NW * 32 * L::ACC_A + NW * (HID + 2);
    static_assert(tiles >= red, "the CTA reduction buffer reuses the tile region");
    constexpr int stage = (grad_use_tma(NA) ? NW * L::ROWS * 3 * NA : 0) + 2 * NW + 8;   // staged rows, mbarriers, TMEM slot
    return round4(param_count(DIN, NOUT)) + tiles + stage + 16;
}

__device__ __forceinline__ f2 pack2u(uint32_t lo, uint32_t hi) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ void unpack2u(f2 v, uint32_t& lo, uint32_t& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}

// One phase-2 pass: TA x 10 register tile per lane (TA = 12 or 8), NG row groups, accumulators parked at TMEM address
// `taddr` between chunks.
template <int TA, int NG, int RS>
__device__ __forceinline__ void tile_pass(const float* __restrict__ wt, int aoff, int doff, int grp, bool busy,
                                          uint32_t taddr) {
    constexpr int NACC = TA * 10;
    f2 acc[NACC / 2];
    {
        uint32_t r[NACC];
        tmem_load<NACC>(taddr, r);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < NACC / 2; ++e) acc[e] = pack2u(r[2 * e], r[2 * e + 1]);
    }
    constexpr int NIT = (64 + NG - 1) / NG;
#pragma unroll 2
    for (int it = 0; it < NIT; ++it) {
        const int row = it * NG + grp;
        if (busy && row < 64) {
            const float* rp = wt + row * RS;
            float a[TA];
#pragma unroll
            for (int q = 0; q < TA / 4; ++q) {
                const float4 v = *reinterpret_cast<const float4*>(rp + aoff + 4 * q);
                a[4 * q] = v.x; a[4 * q + 1] = v.y; a[4 * q + 2] = v.z; a[4 * q + 3] = v.w;
            }
            f2 d[5];
#pragma unroll
            for (int jp = 0; jp < 5; ++jp) d[jp] = *reinterpret_cast<const f2*>(rp + doff + 2 * jp);
#pragma unroll
            for (int ii = 0; ii < TA; ++ii) {
                const f2 aa = pack2(a[ii], a[ii]);
#pragma unroll
                for (int jp = 0; jp < 5; ++jp) acc[ii * 5 + jp] = fma2(aa, d[jp], acc[ii * 5 + jp]);
            }
        }
    }
    {
        uint32_t r[NACC];
#pragma unroll
        for (int e = 0; e < NACC / 2; ++e) unpack2u(acc[e], r[2 * e], r[2 * e + 1]);
        tmem_store<NACC>(taddr, r);
        tmem_wait_st();
    }
}

template <int NA, int DIN, int NOUT, int GRAD_WARPS>
__device__ __forceinline__ void grad_body_v4(const GradParams& P, const rcmarl_grad_job& job, float* smem, int y, int gy) {
    using L = TileLayout4<DIN, NOUT>;
    constexpr int NP = param_count(DIN, NOUT);
    constexpr int R = 2;
    rcmarl_rows Rw = P.rows;
    if (job.time_idx) Rw.time_idx = job.time_idx;
    float* sw = smem;
    float* tiles = smem + round4(NP);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* wt = tiles + warp * (L::ROWS * L::RS);
    constexpr int SROW = 3 * NA;
    constexpr int SWARP = grad_use_tma(NA) ? L::ROWS * SROW : 0;
    float* stage = tiles + GRAD_WARPS * (L::ROWS * L::RS) + warp * SWARP;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + GRAD_WARPS * (L::ROWS * L::RS) + GRAD_WARPS * SWARP);
    uint64_t* bar = bars + warp;
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + GRAD_WARPS);

    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc_all(tslot);                // all 512 columns: one CTA per SM
    pdl_wait();
    stage_weights(sw, job.w, NP);
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    const SmemW W{sw};
    // this warp's TMEM window: lanes 32 * (warp % 4) .. + 31, columns 256 * (warp / 4) .. + 255
    const uint32_t tbase = *tslot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 256);
    const uint32_t tA = tbase, tB = tbase + 128;
    {
        uint32_t z[L::ACC_A];
#pragma unroll
        for (int e = 0; e < L::ACC_A; ++e) z[e] = 0u;
        tmem_store<L::ACC_A>(tA, z);
        tmem_store<L::ACC_B>(tB, z);
        tmem_wait_st();
    }
    // constant columns of this lane's two tile rows, written once
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float* rowp = wt + (lane + 32 * r) * L::RS;
#pragma unroll
        for (int q = 0; q < L::LA1 / 4; ++q)
            if (4 * q >= DIN) st4(rowp + L::OA1 + 4 * q, 4 * q == DIN ? 1.f : 0.f, 0.f, 0.f, 0.f);
        st4(rowp + L::OA2 + 20, 1.f, 0.f, 0.f, 0.f);
        if constexpr (L::L3T) {
            st4(rowp + L::OA3 + 20, 1.f, 0.f, 0.f, 0.f);
            st4(rowp + L::OD3 + 8, 0.f, 0.f, 0.f, 0.f);
        }
    }
    __syncwarp();

    const bool from_ns = (DIN == 2 * NA) && job.kind == RCMARL_IN_NS;
    const int rowf = from_ns ? 2 * NA : 3 * NA;
    const float* in_base = from_ns ? Rw.ns : Rw.sa;
    const bool gather_ok = grad_use_tma(NA) && ((Rw.time_idx == nullptr) || (Rw.n_envs % L::ROWS == 0));
    auto stage_src = [&](int64_t c, const float*& src) -> bool {
        if (!gather_ok || (c + 1) * L::ROWS > Rw.n_rows) return false;
        src = in_base + row_of(Rw, c * L::ROWS) * rowf;
        return (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    };
    uint32_t phase = 0;
    bool staged = false;
    const int64_t cstep = (int64_t)gy * GRAD_WARPS;
#if RCMARL_CHUNK_WARP_MAJOR
    const int64_t cfirst = (int64_t)warp * gy + y;
#else
    const int64_t cfirst = (int64_t)y * GRAD_WARPS + warp;
#endif
    {
        const float* src = nullptr;
        if (cfirst * L::ROWS < Rw.n_rows) staged = stage_src(cfirst, src);
        if (staged && lane == 0) bulk_load(stage, src, (uint32_t)(L::ROWS * rowf * sizeof(float)), bar);
    }

    // phase-2 assignment of this lane
    const bool busyA = lane < L::NGA * L::NTA, busyB = lane < L::NGB * L::NTB;
    const int tileA = busyA ? lane % L::NTA : 0, grpA = busyA ? lane / L::NTA : 0;
    const int tileB = busyB ? lane % L::NTB : 0, grpB = busyB ? lane / L::NTB : 0;
    const int aoffA = L::a_off_A(tileA), doffA = L::d_off_A(tileA);
    const int aoffB = L::a_off_B(tileB), doffB = L::d_off_B(tileB);
    float g3[L::L3T ? 1 : HID + 1];       // scalar nets: output-layer gradient per lane [W3(20) | b3]
#pragma unroll
    for (int j = 0; j < (L::L3T ? 1 : HID + 1); ++j) g3[j] = 0.f;
    float loss = 0.f;

    const int64_t nchunks = (Rw.n_rows + L::ROWS - 1) / L::ROWS;
    for (int64_t c = cfirst; c < nchunks; c += cstep) {
        // ---------------- phase 1: two rows per lane (lane, lane + 32 of the chunk) ----------------
        {
            bool live[R];
            int64_t row[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int64_t m = c * L::ROWS + lane + 32 * r;
                live[r] = m < Rw.n_rows;
                row[r] = row_of(Rw, live[r] ? m : 0);      // dead rows read row 0 and contribute zeros
            }
            float h1[R][HID], h2[R][HID];
            {
                float x[R][DIN];
                if (staged) {
                    mbar_wait(bar, phase);
                    phase ^= 1u;
                    const int skip = (DIN == 2 * NA && !from_ns) ? 1 : 0;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const float* sp = stage + (lane + 32 * r) * rowf;
#pragma unroll
                        for (int k = 0; k < DIN; ++k) x[r][k] = sp[k + skip * (k >> 1)];
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < R; ++r) load_x<NA, DIN>(Rw, job.kind, row[r], x[r]);
                }
                __syncwarp();
                {
                    const int64_t c2 = c + cstep;
                    const float* src = nullptr;
                    staged = (c2 < nchunks) && stage_src(c2, src);
                    if (staged && lane == 0) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        bulk_load(stage, src, (uint32_t)(L::ROWS * rowf * sizeof(float)), bar);
                    }
                }
                dense20_rows<DIN, R>(W, 0, off_b1(DIN), x, h1);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    float* a1 = wt + (lane + 32 * r) * L::RS + L::OA1;
#pragma unroll
                    for (int q = 0; q < L::LA1 / 4; ++q) {
                        if (4 * q >= DIN) continue;                 // constant columns: written once
                        float v[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int k = 4 * q + u;
                            v[u] = k < DIN ? x[r][k < DIN ? k : 0] : (k == DIN ? 1.f : 0.f);
                        }
                        st4(a1 + 4 * q, v[0], v[1], v[2], v[3]);
                    }
                }
            }
            dense20_rows<HID, R>(W, off_W2(DIN), off_b2(DIN), h1, h2);
            float d2[R][HID];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float* rowp = wt + (lane + 32 * r) * L::RS;
#pragma unroll
                for (int q = 0; q < 5; ++q) st4(rowp + L::OA2 + 4 * q, h1[r][4 * q], h1[r][4 * q + 1], h1[r][4 * q + 2], h1[r][4 * q + 3]);
                const float tgt = live[r] ? __ldg(job.target + row[r] * job.target_stride) : 0.f;
                if constexpr (NOUT == 1) {
                    const float e = live[r] ? head1_w<DIN>(W, h2[r]) - tgt : 0.f;
                    loss = fmaf(e, e, loss);
#pragma unroll
                    for (int j = 0; j < HID; ++j) {
                        g3[j] = fmaf(h2[r][j], e, g3[j]);
                        d2[r][j] = W.s(off_W3(DIN) + j) * e * lrelu_grad_from_out(h2[r][j]);
                    }
                    g3[HID] += e;
                } else {
                    float p[NACT], mx, lse, g[NACT];
                    head5_w<DIN>(W, h2[r], p);
                    const int a = (int)__ldg(Rw.sa + row[r] * (3 * NA) + 3 * job.action_agent + 2);
                    float la = 0.f;
#pragma unroll
                    for (int o = 0; o < NACT; ++o) la = (o == a) ? p[o] : la;
                    softmax5(p, mx, lse);
                    loss = fmaf(tgt, (mx + lse) - la, loss);
#pragma unroll
                    for (int o = 0; o < NACT; ++o) g[o] = (p[o] - (o == a ? 1.f : 0.f)) * tgt;
#pragma unroll
                    for (int q = 0; q < 5; ++q) st4(rowp + L::OA3 + 4 * q, h2[r][4 * q], h2[r][4 * q + 1], h2[r][4 * q + 2], h2[r][4 * q + 3]);
                    st4(rowp + L::OD3, g[0], g[1], g[2], g[3]);
                    st4(rowp + L::OD3 + 4, g[4], 0.f, 0.f, 0.f);
#pragma unroll
                    for (int j = 0; j < HID; ++j) {
                        float s = 0.f;
#pragma unroll
                        for (int o = 0; o < NACT; ++o) s = fmaf(W.s(off_W3(DIN) + j * NACT + o), g[o], s);
                        d2[r][j] = s * lrelu_grad_from_out(h2[r][j]);
                    }
                }
#pragma unroll
                for (int q = 0; q < 5; ++q) st4(rowp + L::OD2 + 4 * q, d2[r][4 * q], d2[r][4 * q + 1], d2[r][4 * q + 2], d2[r][4 * q + 3]);
            }
            f2 d2p[R][HID / 2];
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int j = 0; j < HID / 2; ++j) d2p[r][j] = pack2(d2[r][2 * j], d2[r][2 * j + 1]);
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                float d1[R][4];
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) {
                    const int i = 4 * q + ii;
                    f2 s[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) s[r] = pack2(0.f, 0.f);
#pragma unroll
                    for (int qq = 0; qq < 5; ++qq) {
                        const float4 v = W.q(off_W2(DIN) + i * HID + 4 * qq);
                        const f2 w0 = pack2(v.x, v.y), w1 = pack2(v.z, v.w);
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            s[r] = fma2(w0, d2p[r][2 * qq], s[r]);
                            s[r] = fma2(w1, d2p[r][2 * qq + 1], s[r]);
                        }
                    }
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        float se, so;
                        unpack2(s[r], se, so);
                        d1[r][ii] = (se + so) * lrelu_grad_from_out(h1[r][i]);
                    }
                }
#pragma unroll
                for (int r = 0; r < R; ++r)
                    st4(wt + (lane + 32 * r) * L::RS + L::OD1 + 4 * q, d1[r][0], d1[r][1], d1[r][2], d1[r][3]);
            }
        }
        __syncwarp();
        // ---------------- phase 2: exact-fit register tiles, accumulators fetched from / returned to TMEM ----------------
        tile_pass<12, L::NGA, L::RS>(wt, aoffA, doffA, grpA, busyA, tA);
        tile_pass<8, L::NGB, L::RS>(wt, aoffB, doffB, grpB, busyB, tB);
        __syncwarp();
    }

    // ---------------- CTA reduction (fixed order => bitwise reproducible) ----------------
#if RCMARL_PDL_REDUCE
    pdl_launch_dependents();
#endif
    __syncthreads();
    float* red = tiles;                               // [GRAD_WARPS][32][ACC_A] then [GRAD_WARPS][32][ACC_B]
    float* red3 = red + GRAD_WARPS * 32 * L::ACC_A;   // [GRAD_WARPS][HID + 2]: lane-private layer-3 sums + loss
    float* out = P.partial + (int64_t)blockIdx.x * P.stride;
    {
        uint32_t r[L::ACC_A];
        tmem_load<L::ACC_A>(tA, r);
        tmem_wait_ld();
        float4* dst = reinterpret_cast<float4*>(red + (warp * 32 + lane) * L::ACC_A);
#pragma unroll
        for (int q = 0; q < L::ACC_A / 4; ++q)
            dst[q] = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                                 __uint_as_float(r[4 * q + 3]));
    }
    loss = warp_sum(loss);
    if (lane == 0) red3[warp * (HID + 2) + HID + 1] = loss;
    if constexpr (!L::L3T) {
#pragma unroll
        for (int j = 0; j <= HID; ++j) {
            const float s = warp_sum(g3[j]);
            if (lane == 0) red3[warp * (HID + 2) + j] = s;
        }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < L::NTA * L::ACC_A; q += blockDim.x) {
        const int t = q / L::ACC_A, e = q % L::ACC_A;
        const int idx = L::param_A(t, e / 10, e % 10);
        if (idx >= 0) {
            float s = 0.f;
            for (int w = 0; w < GRAD_WARPS; ++w)
#pragma unroll
                for (int g = 0; g < L::NGA; ++g) s += red[(w * 32 + g * L::NTA + t) * L::ACC_A + e];
            out[idx] = s;
        }
    }
    if constexpr (!L::L3T) {
        if (threadIdx.x <= HID) {
            float s = 0.f;
            for (int w = 0; w < GRAD_WARPS; ++w) s += red3[w * (HID + 2) + threadIdx.x];
            out[(threadIdx.x < HID ? off_W3(DIN) : off_b3(DIN, 1) - HID) + threadIdx.x] = s;
        }
    }
    if (threadIdx.x == 32) {
        float s = 0.f;
        for (int w = 0; w < GRAD_WARPS; ++w) s += red3[w * (HID + 2) + HID + 1];
        out[NP] = s;
    }
    __syncthreads();
    {
        uint32_t r[L::ACC_B];
        tmem_load<L::ACC_B>(tB, r);
        tmem_wait_ld();
        float4* dst = reinterpret_cast<float4*>(red + (warp * 32 + lane) * L::ACC_B);
#pragma unroll
        for (int q = 0; q < L::ACC_B / 4; ++q)
            dst[q] = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                                 __uint_as_float(r[4 * q + 3]));
    }
    tmem_fence_before_sync();
    __syncthreads();
    for (int q = threadIdx.x; q < L::NTB * L::ACC_B; q += blockDim.x) {
        const int t = q / L::ACC_B, e = q % L::ACC_B;
        const int idx = L::param_B(t, e / 10, e % 10);
        if (idx >= 0) {
            float s = 0.f;
            for (int w = 0; w < GRAD_WARPS; ++w)
#pragma unroll
                for (int g = 0; g < L::NGB; ++g) s += red[(w * 32 + g * L::NTB + t) * L::ACC_B + e];
            out[idx] = s;
        }
    }
    if (warp == 0) tmem_dealloc_all(*tslot);
}

template <int NA, int LOSS>
__global__ void __launch_bounds__(32 * grad4_warps<NA, LOSS>(), 1) grad_kernel_v4(const __grid_constant__ GradParams P) {
    extern __shared__ __align__(16) float smem[];
    constexpr int NW = grad4_warps<NA, LOSS>();
    int j = 0;
    while (j + 1 < P.n_jobs && (int)blockIdx.x >= P.cta_first[j + 1]) ++j;
    const rcmarl_grad_job& job = P.jobs[j];
    const int y = (int)blockIdx.x - P.cta_first[j], gy = P.cta_first[j + 1] - P.cta_first[j];
    if (LOSS == RCMARL_LOSS_CE) {
        grad_body_v4<NA, 2 * NA, NACT, NW>(P, job, smem, y, gy);
    } else if (job.kind == RCMARL_IN_SA) {
        grad_body_v4<NA, 3 * NA, 1, NW>(P, job, smem, y, gy);
    } else {
        grad_body_v4<NA, 2 * NA, 1, NW>(P, job, smem, y, gy);
    }
}

}  // namespace rcmarl
