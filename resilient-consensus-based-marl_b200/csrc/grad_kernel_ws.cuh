// grad_kernel_ws.cuh -- warp-specialised fused forward + backward of the scalar-output nets (critic / team reward) at
// n_agents = 5: tensor cores (tcgen05, 3xTF32) for the three dense products of a row, CUDA cores for the weight-gradient
// outer products, in different warps of one CTA (replaces the fwd/bwd inside critic.fit / TR.fit,
// agents/resilient_CAC_agents.py:118,136 and agents/adversarial_CAC_agents.py:133,150,163).
//
// Why: grad_kernel (grad_kernel.cuh) runs both phases in every warp.  Its phase 1 feeds the FMA pipes from shared memory
// (2 wavefronts per weight quad) and keeps 64 phase-2 accumulators alive meanwhile, so neither the FMA pipes (60 %) nor
// the shared-memory pipe (74 %) saturate (profiles/r01_final_ncu.md).  The first tensor-core version (round 1, removed)
// moved the dense products to tcgen05 but kept both phases in one thread: each 128-row tile then waits for three MMA round
// trips in sequence and ran SLOWER (12.1 vs 10.0 ms).  Here the two halves are decoupled:
//
//   producer groups (2 x 4 warps, thread = buffer row = TMEM lane, 128 registers)
//       z1 = [x | 1] . [W1; b1]   ->  h1 = lrelu(z1)          tcgen05.mma, A operand written to TMEM by the row's thread as
//       z2 = [h1 | 1] . [W2; b2]  ->  h2, out, e, delta2       tf32 hi / lo halves, B = split weights in shared memory,
//       u  = delta2 . W2^T        ->  delta1 = u * lrelu'(h1)  accumulator read back with tcgen05.ld
//     and write the row's [x | h1 | delta1 | delta2] into a shared-memory tile buffer (two 128-row buffers per group); the
//     output-layer gradient (21 values) and the loss are accumulated per thread.  A tile costs two MMA round trips, not
//     three: the next tile's first layer is issued together with this tile's backward product (ws_produce).
//   consumer warps (2 teams x 4 warps, two per scheduler, 128 registers)
//       sum over the buffer's rows of the outer products [x | 1] (x) delta1 and [h1 | 1] (x) delta2 as 8 x 10 register
//       tiles (5 a-tiles x 2 delta halves x 3 row groups = 30 lanes per warp; 5 LDS.128 per 40 FFMA2), accumulated across
//       ALL tiles of the CTA.
//   Buffers are handed over with full / empty mbarriers; tile q goes to producer group q % 2, that group's ring buffer
//   (q / 2) % 2 and consumer team q % 2 (static, so every sum keeps a fixed order: results are bitwise reproducible).
// While a producer group waits for its MMAs the other group and the consumers own the issue slots; the MMAs themselves
// run beside the FMA pipes.
#pragma once
#include "tc_common.cuh"

namespace rcmarl {

#ifndef RCMARL_WS_GROUPS
#define RCMARL_WS_GROUPS 2
#endif
constexpr int WS_GROUPS = RCMARL_WS_GROUPS;                   // producer groups of 4 warps (2 or 3)
constexpr int WS_CONS = 8;                                    // consumer warps (two per scheduler), in two teams of four
constexpr int WS_WARPS = 4 * WS_GROUPS + WS_CONS;             // 16
constexpr int WS_THREADS = 32 * WS_WARPS;                     // 512
constexpr int WS_RING = 2;                                    // tile buffers per stream (producer group g <-> consumer team g)
constexpr int WS_NBUF = WS_RING * WS_GROUPS;                  // 4: each stream owns its own ring -- the uses of a buffer are then
                                                              // strictly ordered by ONE producer / consumer pair, which the
                                                              // parity waits of the mbarriers need (a shared ring let one stream
                                                              // run two uses ahead of the other and alias the parity)
// Buffer row = 23 x 16 bytes: the five 8-float a-tiles ([x | 1 0..] as two, [h1 | 1 0 0 0] as three) and the four 12-float delta
// halves (delta1 / delta2 as 10 values + 2 zeros each), placed so that EVERY consumer LDS.128 is conflict-free: within a
// quarter-warp the lanes (a-tile, delta half, row group) hit eight different 16-byte bank groups or the same address (found by
// exhaustive search, tools/ws_bank_layout.py; the straightforward order x | h1 | delta1 | delta2 cost 26 wavefronts per step
// instead of 20, and the shared-memory pipe was 64 % busy).  23 is odd, so the producers' row-per-thread STS.128 spread too.
constexpr int WS_ROWF = 92;
constexpr int WS_A0 = 32, WS_A1 = 76, WS_A2 = 84, WS_A3 = 68, WS_A4 = 24;   // float offsets of the a-tiles: x[0..7], x[8..15], h1[0..7], h1[8..15], [h1[16..19] 1 0 0 0]
constexpr int WS_D1A = 40, WS_D1B = 12, WS_D2A = 56, WS_D2B = 0;            // delta1 halves, delta2 halves (float 52..55 is padding)
constexpr int WS_TILE_ROWS = 128;
constexpr int WS_NG = 3;                                      // row groups per consumer warp (5 a-tiles x 2 delta halves x 3 = 30 lanes)
constexpr int WS_TEAM = 4;                                    // consumer warps that share one tile
constexpr int WS_ACC = 40;                                    // packed accumulator pairs per lane: 8 (a) x 10 (half of delta)
// 16 warps x 32 lanes x 128 registers = the whole register file: no setmaxnreg needed (the first versions ran 4 consumer
// warps with 8 x 20 tiles at 240 registers; one FFMA2 stream per scheduler issued only ~30 % of the cycles, and the consumers
// were the bottleneck -- two narrower consumer warps per scheduler cover each other's latencies)
// back-off of a producer group waiting for its MMAs: first sleep after the layer-2 batch (9 MMAs) / after the merged backward +
// next-layer-1 batch (15 MMAs), then the poll interval, in ns (a sweep of 0 .. 450 ns changed nothing, profiles/r02_kernel_experiments.md)
#ifndef RCMARL_WS_SPLIT_TRUNC
#define RCMARL_WS_SPLIT_TRUNC 1
#endif
#ifndef RCMARL_WS_SLEEP2
#define RCMARL_WS_SLEEP2 120
#endif
#ifndef RCMARL_WS_SLEEP3
#define RCMARL_WS_SLEEP3 120
#endif
#ifndef RCMARL_WS_POLL
#define RCMARL_WS_POLL 40
#endif
// Which producer bookkeeping runs in the shadow of the tile's MMA batches instead of ahead of them (bit 0: buffer claim + x / h1
// filing and bit 1: the next tile's loads, behind the layer-2 issue; bit 2: delta2 filing behind the backward issue).  Measured
// per regime (profiles/r02_kernel_experiments.md): the long sweeps of rcmarl_grad are fastest with everything ahead (7.91 vs
// 8.31 ms), the 10-tile sweeps of a mini-batch step with everything deferred (47.4 vs 50.3 us per step).
#ifndef RCMARL_WS_SHADOW_SWEEP
#define RCMARL_WS_SHADOW_SWEEP 0
#endif
#ifndef RCMARL_WS_SHADOW_STEP
#define RCMARL_WS_SHADOW_STEP 6
#endif
constexpr int WS_SHADOW_SWEEP = RCMARL_WS_SHADOW_SWEEP, WS_SHADOW_STEP = RCMARL_WS_SHADOW_STEP;
#ifndef RCMARL_WS_POLL_EMPTY                                   // producers waiting for a free buffer / consumers for a full one
#define RCMARL_WS_POLL_EMPTY 200
#endif
#ifndef RCMARL_WS_POLL_FULL
#define RCMARL_WS_POLL_FULL 100
#endif

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// mbarrier wait with back-off.  try_wait returns after a few tens of nanoseconds whether or not the phase completed, so a
// plain poll loop is ~6 instructions per iteration per waiting warp; in the first version a third of all issued instructions
// were wait loops, and they competed for issue slots with the consumer warp of the same scheduler (the FMA pipes were 45 %
// busy while the producers waited for free buffers).  Sleeping between polls hands those slots to the warps that compute.
// A barrier that never completes ends in a trap (~1 s), not a hang.
template <int FIRST_NS, int NS>
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0, polls = 0;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return;
    if (FIRST_NS > 0) __nanosleep(FIRST_NS);
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
        __nanosleep(NS);
        if (++polls > (1u << 23)) __trap();
    }
}
// one lane of a converged warp (elect.sync): unlike `lane == 0`, ptxas knows the guarded region runs single-threaded and emits the
// tcgen05 instructions (uniform-register operands) without a per-instruction election loop around each of them
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void named_barrier(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// t * LeakyReLU'(pre-activation), decided on the layer's OUTPUT h (same sign): a compare and a predicated multiply instead of
// compare + select + multiply (40 of these per buffer row)
__device__ __forceinline__ float ws_times_lrelu_grad(float t, float h) {
    asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %1, 0f00000000;\n\t@!p mul.f32 %0, %0, %2;\n\t}" : "+f"(t) : "f"(h), "f"(SLOPE));
    return t;
}

// tf32 hi / lo halves of this thread's K operand values -> TMEM columns [col_hi, col_hi + K) and [col_lo, col_lo + K) of its
// lane, 8 columns at a time (16 live temporaries instead of 2 K: the producers run on a 128-register budget)
template <int K>
__device__ __forceinline__ void ws_store_operand(uint32_t tlane, int col_hi, int col_lo, const float (&v)[K]) {
    static_assert(K % 8 == 0, "operand width");
#pragma unroll
    for (int c = 0; c < K / 8; ++c) {
        uint32_t h[8], l[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            // hi = the value truncated to tf32 (one LOP; the tensor core ignores the 13 low mantissa bits anyway), lo = value - hi
            // (exact).  Rounding hi to nearest (+ 0x1000 first) halves |lo| but costs a third operation on each of the 64 operand
            // elements of a row: one product is within 2.2e-7 of fp64 either way (tools/experiments/tf32x3_error.py), the
            // gradient sweep 3 % faster (profiles/r02_kernel_experiments.md).  -DRCMARL_WS_SPLIT_TRUNC=0 restores the rounding.
#if RCMARL_WS_SPLIT_TRUNC
            h[k] = __float_as_uint(v[8 * c + k]) & 0xFFFFE000u;
#else
            h[k] = (__float_as_uint(v[8 * c + k]) + 0x1000u) & 0xFFFFE000u;
#endif
            l[k] = __float_as_uint(v[8 * c + k] - __uint_as_float(h[k]));
        }
        tmem_st8(tlane + col_hi + 8 * c, h);
        tmem_st8(tlane + col_lo + 8 * c, l);
    }
}

constexpr int ws_smem_floats(int n_slots = 1) {
    // packed net(s) | alignment slack | B operands (hi + lo of 32x16, 32x24, 32x24) | tile buffers | barriers + tmem slot
    return n_slots * round32(param_count(15, 1)) + 32 + 2 * TC_N * 16 + 4 * TC_N * 24 + WS_NBUF * WS_TILE_ROWS * WS_ROWF + 256;
}

// packed-parameter index of element (ii, j) of consumer tile t (a rows 8 t .. 8 t + 7 of [x | 1] for t < 2, of [h1 | 1] else)
template <int DIN>
__device__ __forceinline__ int ws_tile_param(int t, int ii, int j) {
    if (t < 2) {
        const int i = 8 * t + ii;
        return i < DIN ? i * HID + j : (i == DIN ? off_b1(DIN) + j : -1);
    }
    const int i = 8 * (t - 2) + ii;
    return i < HID ? off_W2(DIN) + i * HID + j : (i == HID ? off_b2(DIN) + j : -1);
}

// consumer: acc[8][10] += a (x) delta-half for one buffer row (5 LDS, 40 FFMA2 with a broadcast scalar operand); MASKED: the
// a-values are scaled by `keep` (0 for a row beyond the tile)
template <bool MASKED>
__device__ __forceinline__ void ws_consume_row(const float* __restrict__ rp, int acol, int dcol, float keep, f2 (&acc)[WS_ACC]) {
    const float4 a0 = *reinterpret_cast<const float4*>(rp + acol);
    const float4 a1 = *reinterpret_cast<const float4*>(rp + acol + 4);
    const float4 d0 = *reinterpret_cast<const float4*>(rp + dcol);
    const float4 d1 = *reinterpret_cast<const float4*>(rp + dcol + 4);
    const float4 d2 = *reinterpret_cast<const float4*>(rp + dcol + 8);
    const f2 d[5] = {pack2(d0.x, d0.y), pack2(d0.z, d0.w), pack2(d1.x, d1.y), pack2(d1.z, d1.w), pack2(d2.x, d2.y)};
    float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    if constexpr (MASKED) {
#pragma unroll
        for (int ii = 0; ii < 8; ++ii) a[ii] *= keep;
    }
#pragma unroll
    for (int ii = 0; ii < 8; ++ii) {
        const f2 aa = pack2(a[ii], a[ii]);
#pragma unroll
        for (int jp = 0; jp < 5; ++jp) fma2_acc(acc[ii * 5 + jp], aa, d[jp]);
    }
}

// a 20-vector as two halves of 10 values + 2 zeros each (the consumers' delta operands)
template <int N>
__device__ __forceinline__ void ws_store_halves(float* pa, float* pb, const float (&d)[N]) {
    st4(pa, d[0], d[1], d[2], d[3]);
    st4(pa + 4, d[4], d[5], d[6], d[7]);
    st4(pa + 8, d[8], d[9], 0.f, 0.f);
    st4(pb, d[10], d[11], d[12], d[13]);
    st4(pb + 4, d[14], d[15], d[16], d[17]);
    st4(pb + 8, d[18], d[19], 0.f, 0.f);
}
// this tile's features and first hidden layer -> the row's a-tiles
template <int DIN>
__device__ __forceinline__ void ws_file_inputs(float* rowp, const float (&xr)[DIN], const float (&h1)[HID]) {
    st4(rowp + WS_A0, xr[0], xr[1], xr[2], xr[3]);
    st4(rowp + WS_A0 + 4, xr[4], xr[5], xr[6], xr[7]);
    if constexpr (DIN == 15) {
        st4(rowp + WS_A1, xr[8], xr[9], xr[10], xr[11]);
        st4(rowp + WS_A1 + 4, xr[12], xr[13], xr[14], 1.f);
    } else {
        st4(rowp + WS_A1, xr[8], xr[9], 1.f, 0.f);
        st4(rowp + WS_A1 + 4, 0.f, 0.f, 0.f, 0.f);
    }
    st4(rowp + WS_A2, h1[0], h1[1], h1[2], h1[3]);
    st4(rowp + WS_A2 + 4, h1[4], h1[5], h1[6], h1[7]);
    st4(rowp + WS_A3, h1[8], h1[9], h1[10], h1[11]);
    st4(rowp + WS_A3 + 4, h1[12], h1[13], h1[14], h1[15]);
    st4(rowp + WS_A4, h1[16], h1[17], h1[18], h1[19]);
}

// inputs of row r of tile q of this CTA: features, target, and whether the row exists
template <int NA, int DIN>
__device__ __forceinline__ void ws_fetch(const rcmarl_rows& Rw, const rcmarl_grad_job& job, int y, int gy, int q, int r,
                                         float (&xr)[DIN], float& tgt, bool& live) {
    const int64_t m = ((int64_t)y + (int64_t)q * gy) * WS_TILE_ROWS + r;
    live = m < Rw.n_rows;
    const int64_t row = row_of(Rw, live ? m : 0);
    load_x<NA, DIN>(Rw, job.kind, row, xr);
    tgt = live ? __ldg(job.target + row * job.target_stride) : 0.f;
}

// ---- shared-memory layout (the same for both nets: the parameter region is sized for the larger one) ----
struct WsShared {
    float *sw, *b1h, *b1l, *b2h, *b2l, *b3h, *b3l, *bufs, *red3p;
    uint64_t *full, *empty, *mma_bar;
    uint32_t* tslot;
};
constexpr int WS_SLOT = round32(param_count(15, 1));                // floats per parameter slot (768)
__device__ __forceinline__ WsShared ws_carve(float* smem, int n_slots = 1) {
    const int NPMAX = (n_slots - 1) * WS_SLOT + param_count(15, 1);
    WsShared S;
    S.sw = smem;
    // 128-byte alignment by OFFSET arithmetic on the shared-memory pointer: a round trip through uintptr_t would make every
    // later access a generic LD / ST (L1TEX path, long scoreboard) instead of LDS / STS -- measured in the first version
    const uint32_t sbase = smem_u32(smem + NPMAX);
    S.b1h = smem + NPMAX + (((sbase + 127u) & ~127u) - sbase) / 4u;
    S.b1l = S.b1h + TC_N * 16;
    S.b2h = S.b1l + TC_N * 16;
    S.b2l = S.b2h + TC_N * 24;
    S.b3h = S.b2l + TC_N * 24;
    S.b3l = S.b3h + TC_N * 24;
    S.bufs = S.b3l + TC_N * 24;                                       // [WS_NBUF][128][WS_ROWF]
    uint64_t* bars = reinterpret_cast<uint64_t*>(S.bufs + WS_NBUF * WS_TILE_ROWS * WS_ROWF);
    S.full = bars;                                                    // [WS_NBUF]  producers -> consumer (4 warp arrivals)
    S.empty = bars + WS_NBUF;                                         // [WS_NBUF]  consumer team -> producers (WS_TEAM arrivals)
    S.mma_bar = bars + 2 * WS_NBUF;                                   // [WS_GROUPS]
    S.tslot = reinterpret_cast<uint32_t*>(bars + 2 * WS_NBUF + WS_GROUPS);
    S.red3p = reinterpret_cast<float*>(bars) + 32;                    // [8 producer warps][HID + 2]: their parked sums (own region)
    return S;
}
// barriers + tensor memory; touches no global memory (may run before pdl_wait)
__device__ __forceinline__ void ws_init(const WsShared& S) {
    if (threadIdx.x == 0) {
        for (int b = 0; b < WS_NBUF; ++b) { mbar_init(S.full + b, 4); mbar_init(S.empty + b, WS_TEAM); }
        for (int g = 0; g < WS_GROUPS; ++g) mbar_init(S.mma_bar + g, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if ((threadIdx.x >> 5) == 0) tmem_alloc_all(S.tslot);
}
// the constant column block [1 0 0 0] behind h1 of every buffer row (the bias row of the layer-2 weight gradient): written
// once; the scratch use of the buffers at the end of a sweep overwrites it, so ws_pads() is repeated after every reduction
__device__ __forceinline__ void ws_pads(const WsShared& S) {
    for (int i = threadIdx.x; i < WS_NBUF * WS_TILE_ROWS; i += blockDim.x) st4(S.bufs + (int64_t)i * WS_ROWF + WS_A4 + 4, 1.f, 0.f, 0.f, 0.f);
}

// B operands from the staged parameters (canonical K-major, tf32 hi / lo): B1[n][k] = [W1; b1][k][n],
// B2[n][k] = [W2; b2][k][n], B3[n][k] = W2[n][k].  All threads; followed by ws_operands_visible() + a CTA-wide barrier.
template <int DIN>
__device__ __forceinline__ void ws_build_operands(const WsShared& S) {
    ws_pads(S);
    const float* sw = S.sw;
    for (int i = threadIdx.x; i < TC_N * 16; i += blockDim.x) {
        const int n = i / 16, k = i % 16;
        const float v = n < HID ? (k < DIN ? sw[k * HID + n] : (k == DIN ? sw[off_b1(DIN) + n] : 0.f)) : 0.f;
        const float h = to_tf32(v);
        S.b1h[tc_canon(n, k)] = h;
        S.b1l[tc_canon(n, k)] = v - h;
    }
    for (int i = threadIdx.x; i < TC_N * 24; i += blockDim.x) {
        const int n = i / 24, k = i % 24;
        const float v2 = n < HID ? (k < HID ? sw[off_W2(DIN) + k * HID + n] : (k == HID ? sw[off_b2(DIN) + n] : 0.f)) : 0.f;
        const float v3 = (n < HID && k < HID) ? sw[off_W2(DIN) + n * HID + k] : 0.f;
        const float h2v = to_tf32(v2), h3v = to_tf32(v3);
        S.b2h[tc_canon(n, k)] = h2v;
        S.b2l[tc_canon(n, k)] = v2 - h2v;
        S.b3h[tc_canon(n, k)] = h3v;
        S.b3l[tc_canon(n, k)] = v3 - h3v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic smem writes -> async-proxy (MMA) reads
    tmem_fence_before_sync();
}

// tiles of CTA y of gy for a row set: T = y + q * gy, q = 0 .. nq - 1
__device__ __forceinline__ int ws_tile_count(int64_t n_rows, int y, int gy) {
    const int64_t ntiles = (n_rows + WS_TILE_ROWS - 1) / WS_TILE_ROWS;
    return (int)((ntiles > y) ? (ntiles - y + gy - 1) / gy : 0);
}

// Debug timeline (-DRCMARL_WS_TIMELINE=1, tools/ws_timeline.py): thread 0 of CTA 0 records clock64() at the stage boundaries
// of its first tiles; read back with rcmarl_debug_timeline().  Compiled out of the shipped library.
#ifndef RCMARL_WS_TIMELINE
#define RCMARL_WS_TIMELINE 0
#endif
#if RCMARL_WS_TIMELINE
__device__ long long g_ws_timeline[64 * 16];
__device__ long long g_mb_timeline[64 * 16];                  // persistent mini-batch kernel: [step][stage], CTA 0 / thread 0
#define MB_TICK(k) do { if (mb_tl_on && mb_tl_step < 64) g_mb_timeline[mb_tl_step * 16 + (k)] = clock64(); } while (0)
#define WS_TICK(k) do { if (tl_on && tl_tile < 64) g_ws_timeline[tl_tile * 16 + (k)] = clock64(); } while (0)
#else
#define WS_TICK(k) do { } while (0)
#define MB_TICK(k) do { } while (0)
#endif

// ---- producer: tiles q = group, group + WS_GROUPS, ... of this sweep.  nbase = tiles this stream pushed through its ring in
// earlier sweeps of the same kernel (buffer index and barrier parities continue across sweeps); advanced here. ----
// inputs of a stream's first tile, loaded ahead of the sweep (the persistent mini-batch kernel fetches the next step's first
// rows before the step's reduction / exchange, so the sweep does not start with an exposed DRAM round trip)
template <int DIN>
struct WsFirst {
    float x[DIN];
    float tgt;
    bool live;
};
template <int NA, int DIN, int SHADOW, bool PREFETCHED = false>
__device__ __forceinline__ void ws_produce(const WsShared& S, const rcmarl_rows& Rw, const rcmarl_grad_job& job, int y, int gy,
                                           int nq, uint32_t& nbase, uint32_t& mph, float (&g3)[HID + 1], float& loss,
                                           const WsFirst<DIN>* first = nullptr) {
    static_assert((SHADOW & 3) != 1, "the next tile's loads overwrite the features the deferred filing still needs");
    constexpr int K1 = 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = warp >> 2, gwarp = warp & 3;
    const int r = gwarp * 32 + lane;                                 // row of the tile = TMEM lane
    const uint32_t tmem = *S.tslot + (uint32_t)(group * 256);        // this group's column window
    const uint32_t tlane = tmem + ((uint32_t)(gwarp * 32) << 16);
    // columns: layer-1 operand hi / lo | layer-2 and backward operand hi / lo | layer-1 accumulator | layer-2 / backward acc.
    constexpr int CA1H = 0, CA1L = 16, CA2H = 32, CA2L = 56, CD1 = 80, CD2 = 112;
    const uint32_t idesc = umma_idesc_tf32(128, TC_N);
    uint64_t* mbar = S.mma_bar + group;
    const SmemW W{S.sw};
#if RCMARL_WS_TIMELINE
    const bool tl_on = blockIdx.x == 0 && threadIdx.x == 0;
    int tl_tile = -1;
#endif
    if (group >= nq) return;

    // Schedule of one stream (the tiles q = group, group + 2, ... of this CTA), two MMA round trips per tile instead of three:
    //   prologue   x(q0) -> TMEM, L1(q0)                                   -> h1(q0)
    //   per tile   h1 -> TMEM, L2(q)                                         -> h2, e, delta2
    //              delta2 -> TMEM  and  x(next) -> TMEM,  L3(q) + L1(next)   -> delta1(q) (tile complete), h1(next)
    // The next tile's first layer rides on the current tile's backward product (independent operands, one commit, one wait);
    // the inputs of the tile after that are fetched while those MMAs run.
    float xr[DIN];
    float tgt = 0.f;
    bool live = false;
    float h1[HID];
    if constexpr (PREFETCHED) {
#pragma unroll
        for (int k = 0; k < DIN; ++k) xr[k] = first->x[k];
        tgt = first->tgt;
        live = first->live;
    } else {
        ws_fetch<NA, DIN>(Rw, job, y, gy, group, r, xr, tgt, live);
    }
    {   // ---------------- prologue: layer 1 of the stream's first tile ----------------
        float x[K1];
#pragma unroll
        for (int k = 0; k < DIN; ++k) x[k] = xr[k];
#pragma unroll
        for (int k = DIN; k < K1; ++k) x[k] = (k == DIN) ? 1.f : 0.f;
        ws_store_operand<K1>(tlane, CA1H, CA1L, x);
        tmem_wait_st();
        tmem_fence_before_sync();
        named_barrier(1 + group, 128);
        if (gwarp == 0 && elect_one()) {
            tmem_fence_after_sync();
            tc_issue<K1>(tmem + CD1, tmem + CA1H, tmem + CA1L, S.b1h, S.b1l, idesc);
            umma_commit(mbar);
        }
        mbar_wait_sleep<RCMARL_WS_SLEEP2, RCMARL_WS_POLL>(mbar, mph);
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

    for (int q = group; q < nq; q += WS_GROUPS, ++nbase) {
#if RCMARL_WS_TIMELINE
        ++tl_tile;
#endif
        WS_TICK(0);
        const int b = group * WS_RING + (int)(nbase % WS_RING);       // nbase: tiles this stream has pushed through its ring
        float* rowp = S.bufs + ((int64_t)b * WS_TILE_ROWS + r) * WS_ROWF;
        const bool has_next = q + WS_GROUPS < nq;
        if constexpr ((SHADOW & 1) == 0) {
            mbar_wait_sleep<0, RCMARL_WS_POLL_EMPTY>(S.empty + b, ((nbase / WS_RING) & 1u) ^ 1u);
            ws_file_inputs<DIN>(rowp, xr, h1);
        }
        const float tgt_q = tgt;
        const bool live_q = live;
        if constexpr ((SHADOW & 2) == 0) {
            if (has_next) ws_fetch<NA, DIN>(Rw, job, y, gy, q + WS_GROUPS, r, xr, tgt, live);
        }
        // ---------------- layer 2, output layer, delta2 ----------------
        float d2[24];
        {
            float a2[24];
#pragma unroll
            for (int j = 0; j < HID; ++j) a2[j] = h1[j];
            a2[20] = 1.f; a2[21] = 0.f; a2[22] = 0.f; a2[23] = 0.f;
            ws_store_operand<24>(tlane, CA2H, CA2L, a2);
            tmem_wait_st();
            tmem_fence_before_sync();
            WS_TICK(1);
            named_barrier(1 + group, 128);
            WS_TICK(2);
            if (gwarp == 0 && elect_one()) {
                tmem_fence_after_sync();
                tc_issue<24>(tmem + CD2, tmem + CA2H, tmem + CA2L, S.b2h, S.b2l, idesc);
                umma_commit(mbar);
            }
            // In the shadow of those MMAs: claim the tile buffer (the consumers are done with its previous use), file this tile's
            // features (still in xr: their layer-1 product is done) and h1, then start the next tile's loads -- they have the
            // rest of this round trip and the output-layer arithmetic to arrive.  None of it is on the path to the next issue.
            if constexpr ((SHADOW & 1) != 0) {
            mbar_wait_sleep<0, RCMARL_WS_POLL_EMPTY>(S.empty + b, ((nbase / WS_RING) & 1u) ^ 1u);
            ws_file_inputs<DIN>(rowp, xr, h1);
            }
            if constexpr ((SHADOW & 2) != 0) {
                if (has_next) ws_fetch<NA, DIN>(Rw, job, y, gy, q + WS_GROUPS, r, xr, tgt, live);
            }
            mbar_wait_sleep<RCMARL_WS_SLEEP2, RCMARL_WS_POLL>(mbar, mph);
            WS_TICK(3);
            mph ^= 1u;
            tmem_fence_after_sync();
            uint32_t z[24];
            tmem_load<24>(tlane + CD2, z);
            tmem_wait_ld();
            WS_TICK(4);
            float h2[HID];
#pragma unroll
            for (int j = 0; j < HID; ++j) {
                const float zz = __uint_as_float(z[j]);
                h2[j] = fmaxf(zz, SLOPE * zz);
            }
            // Keras MSE (Appendix A.2): dLoss/dout = 2 (out - y) / B; the 2/B is applied by the caller
            const float e = live_q ? head1_w<DIN>(W, h2) - tgt_q : 0.f;
            loss = fmaf(e, e, loss);
#pragma unroll
            for (int j = 0; j < HID; ++j) {
                g3[j] = fmaf(h2[j], e, g3[j]);
                d2[j] = ws_times_lrelu_grad(W.s(off_W3(DIN) + j) * e, h2[j]);
            }
            g3[HID] += e;
            d2[20] = 0.f; d2[21] = 0.f; d2[22] = 0.f; d2[23] = 0.f;
        }
        // ---------------- backward-data of this tile + layer 1 of the stream's next tile ----------------
        {
            if constexpr ((SHADOW & 4) == 0) ws_store_halves(rowp + WS_D2A, rowp + WS_D2B, d2);
            ws_store_operand<24>(tlane, CA2H, CA2L, d2);
            if (has_next) {                                           // xr holds the next tile's features by now
                float x[K1];
#pragma unroll
                for (int k = 0; k < DIN; ++k) x[k] = xr[k];
#pragma unroll
                for (int k = DIN; k < K1; ++k) x[k] = (k == DIN) ? 1.f : 0.f;
                ws_store_operand<K1>(tlane, CA1H, CA1L, x);
            }
            tmem_wait_st();
            tmem_fence_before_sync();
            WS_TICK(5);
            named_barrier(1 + group, 128);
            WS_TICK(6);
            if (gwarp == 0 && elect_one()) {
                tmem_fence_after_sync();
                tc_issue<24>(tmem + CD2, tmem + CA2H, tmem + CA2L, S.b3h, S.b3l, idesc);
                if (has_next) tc_issue<K1>(tmem + CD1, tmem + CA1H, tmem + CA1L, S.b1h, S.b1l, idesc);
                umma_commit(mbar);
            }
            if constexpr ((SHADOW & 4) != 0) ws_store_halves(rowp + WS_D2A, rowp + WS_D2B, d2);   // in the shadow of the MMAs
            mbar_wait_sleep<RCMARL_WS_SLEEP3, RCMARL_WS_POLL>(mbar, mph);
            WS_TICK(7);
            mph ^= 1u;
            tmem_fence_after_sync();
            uint32_t u[24];
            tmem_load<24>(tlane + CD2, u);
            tmem_wait_ld();
            float d1[HID];
#pragma unroll
            for (int i = 0; i < HID; ++i) d1[i] = ws_times_lrelu_grad(__uint_as_float(u[i]), h1[i]);
            ws_store_halves(rowp + WS_D1A, rowp + WS_D1B, d1);
            __syncwarp();
            if (lane == 0) mbar_arrive(S.full + b);                   // release: this warp's 32 rows of the buffer are complete
            WS_TICK(8);
            if (has_next) {
                uint32_t z[24];
                tmem_load<24>(tlane + CD1, z);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < HID; ++j) {
                    const float zz = __uint_as_float(z[j]);
                    h1[j] = fmaxf(zz, SLOPE * zz);
                }
            }
            tmem_fence_before_sync();       // the next MMAs overwrite the accumulators only after the next group barrier
        }
        WS_TICK(9);
    }
}

// ---- consumer warp cw (team cw / 4, member cw % 4): its team takes every second tile, the four members split the tile's
// 43 three-row steps.  A buffer is held for a quarter of the time one warp would need, and each team alternates between the two
// buffers of its producer group's ring. ----
__device__ __forceinline__ void ws_consume(const WsShared& S, int cw, int nq, uint32_t& nbase, f2 (&acc)[WS_ACC]) {
    const int lane = threadIdx.x & 31;
    const bool active = lane < 10 * WS_NG;
    const int combo = active ? lane % 10 : 0;                         // (a-tile, delta half)
    const int atile = combo % 5, half = combo / 5;
    const int grp = active ? lane / 10 : 0;
    const int acol = atile == 0 ? WS_A0 : atile == 1 ? WS_A1 : atile == 2 ? WS_A2 : atile == 3 ? WS_A3 : WS_A4;
    const int dcol = atile < 2 ? (half ? WS_D1B : WS_D1A) : (half ? WS_D2B : WS_D2A);   // x-tiles pair with delta1, h1-tiles with delta2
    const int team = cw / WS_TEAM, member = cw % WS_TEAM;
    constexpr int FULL_STEPS = WS_TILE_ROWS / WS_NG;                  // 42 (+ rows 126, 127 as step 42)
    static_assert(WS_CONS / WS_TEAM == WS_GROUPS, "one consumer team per producer group");
    for (int q = team; q < nq; q += WS_GROUPS, ++nbase) {
        const int b = team * WS_RING + (int)(nbase % WS_RING);
        const uint32_t use = nbase / WS_RING;
        mbar_wait_sleep<0, RCMARL_WS_POLL_FULL>(S.full + b, use & 1u);
        const float* buf = S.bufs + (int64_t)b * WS_TILE_ROWS * WS_ROWF;
        // member m takes steps m, m + 4, ... of the 43 three-row steps: ten for everybody plus an eleventh whose rows may lie
        // beyond the tile (members 2 and 3: rows 128+), handled branch-free by clamping the row and zeroing its a-values.  All
        // eleven steps are straight-line code: as a loop (or behind a branch) ptxas copied 17 accumulators around per step.
        const float* rp = buf + (member * WS_NG + grp) * WS_ROWF;
#pragma unroll
        for (int k = 0; k < FULL_STEPS / WS_TEAM; ++k) ws_consume_row<false>(rp + k * (WS_TEAM * WS_NG * WS_ROWF), acol, dcol, 1.f, acc);
        {
            const int rowi = ((FULL_STEPS / WS_TEAM) * WS_TEAM + member) * WS_NG + grp;
            const int rowc = rowi < WS_TILE_ROWS ? rowi : WS_TILE_ROWS - 1;
            ws_consume_row<true>(buf + rowc * WS_ROWF, acol, dcol, rowi < WS_TILE_ROWS ? 1.f : 0.f, acc);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(S.empty + b);
    }
}

// ---- end of a sweep.  Producers park their sums in a region of their own as soon as their tiles are produced.  Consumers wait
// for ALL consumers (a 256-thread named barrier inside the consumer branch: the other team may still be reading the buffers
// that the scratch overlays), then park theirs in the idle tile buffers.  The CTA-wide barrier that follows sits BEHIND the
// role branch, at one program location; store(i, v) receives the sums for the packed parameters i = 0 .. NP-1, the loss as NP ----
__device__ __forceinline__ float* ws_red(const WsShared& S) { return S.bufs; }
// floats per parked consumer lane: 80 sums + 4 pad = 21 x 16 bytes (odd), so the lanes' STS.128 spread over the bank groups
// (at 80 floats the eight lanes of a quarter-warp hit two bank groups: 4-way conflicts on every store of the epilogue)
constexpr int WS_PARK = 2 * WS_ACC + 4;
__device__ __forceinline__ float* ws_red3(const WsShared& S) { return S.red3p; }
__device__ __forceinline__ void ws_park_producer(const WsShared& S, const float (&g3)[HID + 1], float loss) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* red3 = ws_red3(S);
    named_barrier(1 + (warp >> 2), 128);      // the group is through with its tiles (and its warps are converged for the shuffles)
    loss = warp_sum(loss);
    if (lane == 0) red3[warp * (HID + 2) + HID + 1] = loss;
#pragma unroll
    for (int j = 0; j <= HID; ++j) {
        const float s = warp_sum(g3[j]);
        if (lane == 0) red3[warp * (HID + 2) + j] = s;
    }
}
__device__ __forceinline__ void ws_park_consumer(const WsShared& S, int cw, const f2 (&acc)[WS_ACC]) {
    const int lane = threadIdx.x & 31;
    float4* dst = reinterpret_cast<float4*>(ws_red(S) + (cw * 32 + lane) * WS_PARK);
#pragma unroll
    for (int e = 0; e < WS_ACC / 2; ++e) {
        float4 v;
        unpack2(acc[2 * e], v.x, v.y);
        unpack2(acc[2 * e + 1], v.z, v.w);
        dst[e] = v;
    }
}
template <int DIN, class ST>
__device__ __forceinline__ void ws_cta_sums(const WsShared& S, ST store) {
    constexpr int NP = param_count(DIN, 1);
    const float* red = ws_red(S);
    const float* red3 = ws_red3(S);
    for (int qi = threadIdx.x; qi < 5 * 160; qi += WS_THREADS) {
        const int t = qi / 160, e = qi % 160;                         // a-tile t, element (ii, j) = (e / 20, e % 20)
        const int ii = e / 20, j = e % 20;
        const int idx = ws_tile_param<DIN>(t, ii, j);
        if (idx >= 0) {
            const int half = j / 10, slot = ii * 10 + (j % 10);       // lane combo = half * 5 + t, accumulator slot in the lane
            float s = 0.f;
            for (int c = 0; c < WS_CONS; ++c)
#pragma unroll
                for (int g = 0; g < WS_NG; ++g) s += red[(c * 32 + g * 10 + half * 5 + t) * WS_PARK + slot];
            store(idx, s);
        }
    }
    if (threadIdx.x <= HID) {
        float s = 0.f;
        for (int w = 0; w < 4 * WS_GROUPS; ++w) s += red3[w * (HID + 2) + threadIdx.x];
        store((threadIdx.x < HID ? off_W3(DIN) : off_b3(DIN, 1) - HID) + threadIdx.x, s);
    }
    if (threadIdx.x == 32) {
        float s = 0.f;
        for (int w = 0; w < 4 * WS_GROUPS; ++w) s += red3[w * (HID + 2) + HID + 1];
        store(NP, s);
    }
}
constexpr int WS_BAR_CONS = 8;                                // named barrier of the 8 consumer warps (after their last tile)

// =====================================================================================================================
// one-shot kernel (rcmarl_grad): one sweep, sums to the CTA's partial slot
// =====================================================================================================================
template <int NA, int DIN>
__device__ __forceinline__ void ws_body(const GradParams& P, const rcmarl_grad_job& job, float* smem, int y, int gy) {
    static_assert(NA == 5, "instantiated for n_agents = 5 (operand widths 16 / 24)");
    constexpr int NP = param_count(DIN, 1);
    rcmarl_rows Rw = P.rows;
    if (job.time_idx) Rw.time_idx = job.time_idx;
    const int warp = threadIdx.x >> 5;
    const WsShared S = ws_carve(smem);
    ws_init(S);
    pdl_wait();
    stage_weights(S.sw, job.w, NP);
    __syncthreads();
    ws_build_operands<DIN>(S);
    __syncthreads();
    tmem_fence_after_sync();
    const int nq = ws_tile_count(Rw.n_rows, y, gy);
    float* out = P.partial + (int64_t)blockIdx.x * P.stride;
    if (warp < 4 * WS_GROUPS) {
        uint32_t mph = 0, nbase = 0;
        float g3[HID + 1];
#pragma unroll
        for (int j = 0; j <= HID; ++j) g3[j] = 0.f;
        float loss = 0.f;
        ws_produce<NA, DIN, WS_SHADOW_SWEEP>(S, Rw, job, y, gy, nq, nbase, mph, g3, loss);
        ws_park_producer(S, g3, loss);
    } else {
        const int cw = warp - 4 * WS_GROUPS;
        f2 acc[WS_ACC];                                               // 8 (a) x 10 (delta half), packed as pairs over the delta index
#pragma unroll
        for (int e = 0; e < WS_ACC; ++e) acc[e] = pack2(0.f, 0.f);
        uint32_t nbase = 0;
        ws_consume(S, cw, nq, nbase, acc);
        named_barrier(WS_BAR_CONS, 32 * WS_CONS);                     // every tile consumed
        ws_park_consumer(S, cw, acc);
    }
#if RCMARL_PDL_REDUCE
    pdl_launch_dependents();
#endif
    __syncthreads();                                                  // (B) scratch complete
    ws_cta_sums<DIN>(S, [out](int i, float v) { out[i] = v; });
    tmem_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc_all(*S.tslot);
}

// scalar-output nets at n_agents = 5 (mean-squared-error jobs); everything else stays on grad_kernel
__global__ void __launch_bounds__(WS_THREADS, 1) grad_kernel_ws(const __grid_constant__ GradParams P) {
    extern __shared__ __align__(16) float smem[];
    int j = 0;
    while (j + 1 < P.n_jobs && (int)blockIdx.x >= P.cta_first[j + 1]) ++j;
    const rcmarl_grad_job& job = P.jobs[j];
    const int y = (int)blockIdx.x - P.cta_first[j], gy = P.cta_first[j + 1] - P.cta_first[j];
    if (job.kind == RCMARL_IN_SA) ws_body<5, 15>(P, job, smem, y, gy);
    else ws_body<5, 10>(P, job, smem, y, gy);
}

}  // namespace rcmarl
