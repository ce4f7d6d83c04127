// minibatch_persist.cuh -- the adversaries' mini-batch fits as ONE persistent kernel per rcmarl_minibatch_sgd call
// (replaces critic.fit / TR.fit(batch_size=32, epochs=10) of agents/adversarial_CAC_agents.py:133,150,163,239,251).
//
// Round 1 ran every SGD step as two launches (grad_kernel + fused reduce / apply): 9 400 sequential steps per update
// round at ~17 us of launch gaps, prologues and a separate reduce grid each.  Here the CTAs stay resident for all
// epochs x mini-batches of a call, each chain's parameters live in the shared memory of its CTAs, and a step is
//   1. sweep this CTA's 64-row chunks of the mini-batch (GradCore, register accumulators),
//   2. fixed-order CTA reduction; the CTA's sums go out as level-1 cells {value, seq} (8-byte stores),
//   3. the CTAs of a chain split the parameters into slices: the owner of a slice polls the level-1 cells of all CTAs
//      of its chain for that slice, adds them in CTA order and publishes level-2 cells -- into its own GPU's buffer, or
//      (data parallel) into EVERY rank's buffer over NVLink peer memory (comm.cuh, same cell format),
//   4. every CTA polls the level-2 cells of all parameters (rank-ordered sum over the ranks) and applies the SGD step to
//      its shared-memory copy of the chain's parameters.
// No grid barrier, no atomics, no launches, no host round trip; every sum has a fixed association, so the result is
// bitwise reproducible and identical on every CTA and every rank.  A sequence number is used exactly once per step, so
// stale cells can never be mistaken for fresh ones; the cell buffers must start zeroed (sequence numbers start at 1).
// Why single-buffered level-1 cells are safe: a CTA writes its level-1 cells of step s+1 only after it has read all
// level-2 cells of step s, which exist only after every slice owner of its chain has read the level-1 cells of step s.
// Level-2 cells are double-buffered by sequence parity for the cross-rank case (comm.cuh).
#pragma once
#include "grad_kernel.cuh"
#include "comm.cuh"
#ifndef RCMARL_GRAD_WS
#define RCMARL_GRAD_WS 1
#endif
#if RCMARL_GRAD_WS
#include "grad_kernel_ws.cuh"
#endif

namespace rcmarl {

struct MbChain {
    float* w;                    // packed parameters, updated in place (read at start, written back at the end)
    const float* target;
    const int32_t* time_idx;     // [epochs][n_times] shuffled time rows
    float* loss_out;             // += loss_coef * sum over the first epoch's steps of sum e^2 (may be null)
    int64_t target_stride;
    float lr, loss_coef;
    int32_t kind, loss_accumulate;
};

struct MbParams {
    rcmarl_rows rows;                       // sa / ns / r, row_begin, n_envs, n_agents (n_rows, time_idx set per step)
    MbChain chains[RCMARL_MAX_JOBS];
    int16_t cta_first[RCMARL_MAX_JOBS + 1];
    int32_t n_chains, epochs, n_times, mb_times, stride;
    uint2* cells1;                          // level 1: [n_ctas][stride]
    uint32_t seq1;                          // level-1 sequence number of the call's first step
    CommDev comm;                           // level 2 (world == 1: cells[0] is a local buffer); comm.seq = first step's
};

__device__ __forceinline__ uint2 poll_cell(const uint2* cell, uint32_t seq, uint32_t* err) {
    uint2 x = ld_cell(cell);
    if (x.y != seq) {
        const long long t0 = clock64();
        do {
            if (clock64() - t0 > 40000000000LL) {      // ~20 s: fail loudly instead of hanging the GPU
                if (err) *err = 2u;
                __threadfence_system();
                __trap();
            }
            __nanosleep(RCMARL_CELL_POLL_NS);           // spinning CTAs must not crowd the writers out of the L2 queues
            x = ld_cell(cell);
        } while (x.y != seq);
    }
    return x;
}

template <int NA, int DIN, int NW>
__device__ __forceinline__ void mb_body(const MbParams& P, const MbChain& ch, int j, float* smem, int y, int gy) {
    using Core = GradCore<NA, DIN, 1, NW>;
    constexpr int NP = Core::NP;
    Core core;
    core.setup(smem);
    pdl_wait();
    stage_weights(core.sw, ch.w, NP);
    __syncthreads();
    core.write_pads();

    rcmarl_grad_job gj;
    gj.w = ch.w; gj.target = ch.target; gj.sums = nullptr; gj.time_idx = nullptr;
    gj.target_stride = ch.target_stride; gj.kind = ch.kind; gj.action_agent = 0;
    rcmarl_rows Rw = P.rows;

    const int cta = blockIdx.x;
    uint2* my1 = P.cells1 + (int64_t)cta * P.stride;
    const uint2* chain1 = P.cells1 + (int64_t)P.cta_first[j] * P.stride;
    // slice of this CTA: entries [sl_begin, sl_end) of the NP + 1 sums (the last entry is the loss)
    const int per = (NP + 1 + gy - 1) / gy;
    const int sl_begin = y * per < NP + 1 ? y * per : NP + 1;
    const int sl_end = sl_begin + per < NP + 1 ? sl_begin + per : NP + 1;
    // level-1 gather geometry: S lanes share an entry (S = power of two >= min(gy, 32)), each adds every S-th CTA
    int S = 1;
    while (S < gy && S < 32) S <<= 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sgroup = lane & (S - 1), e_local = lane / S, e_per_warp = 32 / S;
    const int64_t off2 = (int64_t)j * P.stride;
    const CommDev& comm = P.comm;
    uint32_t seq2 = P.comm.seq;
    const int nb = (P.n_times + P.mb_times - 1) / P.mb_times;
    uint32_t seq1 = P.seq1;
    float loss_acc = 0.f;

    for (int e = 0; e < P.epochs; ++e) {
        for (int b = 0; b < nb; ++b, ++seq1, ++seq2) {
            const int cnt = P.n_times - b * P.mb_times < P.mb_times ? P.n_times - b * P.mb_times : P.mb_times;
            Rw.n_rows = (int64_t)cnt * Rw.n_envs;
            Rw.time_idx = ch.time_idx + (int64_t)e * P.n_times + (int64_t)b * P.mb_times;
            core.zero_acc();
            core.sweep(Rw, gj, y, gy);
            // ---- level 1: this CTA's sums
            const uint32_t s1 = seq1;
            core.cta_reduce([my1, s1](int i, float v) { st_cell(my1 + i, v, s1); });
            // ---- slice owner: CTA-ordered sum over the chain's CTAs, publish level 2
            for (int base = sl_begin + warp * e_per_warp; base < sl_end; base += NW * e_per_warp) {
                const int i = base + e_local;
                float s = 0.f;
                if (i < sl_end) {
                    for (int yy = sgroup; yy < gy; yy += S)
                        s += __uint_as_float(poll_cell(chain1 + (int64_t)yy * P.stride + i, s1, comm.error).x);
                }
                for (int o = 1; o < S; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (i < sl_end && sgroup == 0) comm_push(comm, off2 + i, s, seq2);
            }
            // ---- level 2: every CTA applies the step to its own copy (tile region is free: cta_reduce is done with it)
            const float coef = ch.lr * 2.0f / ((float)Rw.n_rows * (float)comm.world);
            for (int i = threadIdx.x; i <= NP; i += blockDim.x) {
                const float tot = comm_wait_total(comm, off2 + i, seq2);
                if (i < NP) core.sw[i] = core.sw[i] - coef * tot;
                else if (e == 0) loss_acc += ch.loss_coef * tot;
            }
            __syncthreads();                 // new parameters visible to all warps; tile region reusable
            core.write_pads();
        }
    }
    if (y == 0) {
        for (int i = threadIdx.x; i < NP; i += blockDim.x) ch.w[i] = core.sw[i];
        if (threadIdx.x == (NP % blockDim.x) && ch.loss_out) *ch.loss_out = ch.loss_accumulate ? *ch.loss_out + loss_acc : loss_acc;
    }
}

template <int NA>
__global__ void __launch_bounds__(32 * grad_warps<NA, RCMARL_LOSS_MSE>(), 1)
mb_persist_kernel(const __grid_constant__ MbParams P) {
    extern __shared__ __align__(16) float smem[];
    constexpr int NW = grad_warps<NA, RCMARL_LOSS_MSE>();
    pdl_launch_dependents();         // the next kernel may be queued; it waits for this grid with griddepcontrol.wait
    int j = 0;
    while (j + 1 < P.n_chains && (int)blockIdx.x >= P.cta_first[j + 1]) ++j;
    const MbChain& ch = P.chains[j];
    const int y = (int)blockIdx.x - P.cta_first[j], gy = P.cta_first[j + 1] - P.cta_first[j];
    if (ch.kind == RCMARL_IN_SA) mb_body<NA, 3 * NA, NW>(P, ch, j, smem, y, gy);
    else mb_body<NA, 2 * NA, NW>(P, ch, j, smem, y, gy);
}

#if RCMARL_GRAD_WS
// =====================================================================================================================
// The same persistent fit on the warp-specialised tensor-core core (grad_kernel_ws.cuh), n_agents = 5.  Both roles run the
// same step loop with the role branch inside the step; every CTA-wide barrier sits behind that branch (one program location).
// =====================================================================================================================
struct MbStepCtx {
    uint2* my1;                 // this CTA's level-1 cells
    const uint2* chain1;        // level-1 cells of the chain's first CTA
    const CommDev* comm;
    int64_t off2;               // the chain's block in the level-2 cells
    int32_t gy, stride, sl_begin, sl_end, S, sgroup, e_local, e_per_warp;
    float loss_coef;
};

// sums of this step -> level-1 cells -> slice owner's CTA-ordered sum -> level-2 cells -> SGD step on the shared-memory copy
template <int DIN>
__device__ __forceinline__ void mb_step_tail(const WsShared& S, const MbStepCtx& c, uint32_t s1, uint32_t s2, float coef,
                                             bool first_epoch, float& loss_acc, bool mb_tl_on = false, int mb_tl_step = 0) {
    constexpr int NP = param_count(DIN, 1);
    const int warp = threadIdx.x >> 5;
    uint2* my1 = c.my1;
    ws_cta_sums<DIN>(S, [my1, s1](int i, float v) { st_cell(my1 + i, v, s1); });
    MB_TICK(5);
    for (int base = c.sl_begin + warp * c.e_per_warp; base < c.sl_end; base += WS_WARPS * c.e_per_warp) {
        const int i = base + c.e_local;
        float s = 0.f;
        if (i < c.sl_end) {
            for (int yy = c.sgroup; yy < c.gy; yy += c.S)
                s += __uint_as_float(poll_cell(c.chain1 + (int64_t)yy * c.stride + i, s1, c.comm->error).x);
        }
        for (int o = 1; o < c.S; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (i < c.sl_end && c.sgroup == 0) comm_push(*c.comm, c.off2 + i, s, s2);
    }
    MB_TICK(6);
    for (int i = threadIdx.x; i <= NP; i += WS_THREADS) {
        const float tot = comm_wait_total(*c.comm, c.off2 + i, s2);
        if (i < NP) S.sw[i] = S.sw[i] - coef * tot;
        else if (first_epoch) loss_acc += c.loss_coef * tot;
    }
    MB_TICK(7);
}

template <int NA, int DIN>
__device__ __forceinline__ void mb_body_ws(const MbParams& P, const MbChain& ch, int j, float* smem, int y, int gy) {
    constexpr int NP = param_count(DIN, 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const WsShared S = ws_carve(smem);
    ws_init(S);
    pdl_wait();
    stage_weights(S.sw, ch.w, NP);
    __syncthreads();

    rcmarl_grad_job gj;
    gj.w = ch.w; gj.target = ch.target; gj.sums = nullptr; gj.time_idx = nullptr;
    gj.target_stride = ch.target_stride; gj.kind = ch.kind; gj.action_agent = 0;
    rcmarl_rows Rw = P.rows;

    MbStepCtx c;
    c.my1 = P.cells1 + (int64_t)blockIdx.x * P.stride;
    c.chain1 = P.cells1 + (int64_t)P.cta_first[j] * P.stride;
    c.comm = &P.comm;
    c.off2 = (int64_t)j * P.stride;
    c.gy = gy; c.stride = P.stride;
    const int per = (NP + 1 + gy - 1) / gy;
    c.sl_begin = y * per < NP + 1 ? y * per : NP + 1;
    c.sl_end = c.sl_begin + per < NP + 1 ? c.sl_begin + per : NP + 1;
    int Sl = 1;
    while (Sl < gy && Sl < 32) Sl <<= 1;
    c.S = Sl; c.sgroup = lane & (Sl - 1); c.e_local = lane / Sl; c.e_per_warp = 32 / Sl;
    c.loss_coef = ch.loss_coef;
    const int nb = (P.n_times + P.mb_times - 1) / P.mb_times;
    const float world = (float)P.comm.world;
    uint32_t seq1 = P.seq1, seq2 = P.comm.seq, nbase = 0;      // nbase: this stream's tiles so far (ring bookkeeping)
    float loss_acc = 0.f;

    // One step loop for both roles: the role branch sits INSIDE the step, every CTA-wide barrier behind it at one program
    // location (nothing role-specific is live there: both roles have parked their sums by then).
    const bool producer = warp < 4 * WS_GROUPS;
    const int cw = warp - 4 * WS_GROUPS;
    const int group = warp >> 2, prow = (warp & 3) * 32 + lane;      // producer group, row of the tile
    uint32_t mph = 0;
#if RCMARL_WS_TIMELINE
    const bool mb_tl_on = blockIdx.x == 0 && threadIdx.x == 0;       // ticks 0 .. 9: first producer thread
    int mb_tl_step = -1;
#else
    constexpr bool mb_tl_on = false;
    constexpr int mb_tl_step = 0;
#endif
    // rows of step (e, b): a mini-batch is mb_times time rows x all environments of this rank
    auto step_rows = [&](int e, int b, rcmarl_rows& R) {
        const int cnt = P.n_times - b * P.mb_times < P.mb_times ? P.n_times - b * P.mb_times : P.mb_times;
        R.n_rows = (int64_t)cnt * R.n_envs;
        R.time_idx = ch.time_idx + (int64_t)e * P.n_times + (int64_t)b * P.mb_times;
    };
    WsFirst<DIN> nf;                                                  // first tile of the NEXT sweep, fetched one step ahead
    nf.tgt = 0.f; nf.live = false;
#pragma unroll
    for (int k = 0; k < DIN; ++k) nf.x[k] = 0.f;
    step_rows(0, 0, Rw);
    if (producer && group < ws_tile_count(Rw.n_rows, y, gy)) ws_fetch<NA, DIN>(Rw, gj, y, gy, group, prow, nf.x, nf.tgt, nf.live);
    for (int e = 0; e < P.epochs; ++e) {
        for (int b = 0; b < nb; ++b, ++seq1, ++seq2) {
#if RCMARL_WS_TIMELINE
            ++mb_tl_step;
#endif
            MB_TICK(0);
            ws_build_operands<DIN>(S);
            __syncthreads();
            MB_TICK(1);
            tmem_fence_after_sync();
            step_rows(e, b, Rw);
            const int nq = ws_tile_count(Rw.n_rows, y, gy);
            const float n_rows_f = (float)Rw.n_rows;
            if (producer) {
                float g3[HID + 1];
#pragma unroll
                for (int k = 0; k <= HID; ++k) g3[k] = 0.f;
                float loss = 0.f;
                ws_produce<NA, DIN, WS_SHADOW_STEP, true>(S, Rw, gj, y, gy, nq, nbase, mph, g3, loss, &nf);
                {   // the next step's first rows: in flight during this step's reduction, exchange and SGD
                    const int bn = b + 1 < nb ? b + 1 : 0, en = b + 1 < nb ? e : e + 1;
                    if (en < P.epochs) {
                        rcmarl_rows Rn = Rw;
                        step_rows(en, bn, Rn);
                        if (group < ws_tile_count(Rn.n_rows, y, gy)) ws_fetch<NA, DIN>(Rn, gj, y, gy, group, prow, nf.x, nf.tgt, nf.live);
                    }
                }
                MB_TICK(2);
                ws_park_producer(S, g3, loss);
                MB_TICK(3);
            } else {
#if RCMARL_WS_TIMELINE
                const bool mb_tl_on = blockIdx.x == 0 && threadIdx.x == 32 * 4 * WS_GROUPS;   // first consumer thread: ticks 10 .. 12
#endif
                f2 acc[WS_ACC];
#pragma unroll
                for (int k = 0; k < WS_ACC; ++k) acc[k] = pack2(0.f, 0.f);
                ws_consume(S, cw, nq, nbase, acc);
                MB_TICK(10);
                named_barrier(WS_BAR_CONS, 32 * WS_CONS);             // every tile consumed
                MB_TICK(11);
                ws_park_consumer(S, cw, acc);
                MB_TICK(12);
            }
            __syncthreads();                                          // scratch complete
            MB_TICK(4);
            mb_step_tail<DIN>(S, c, seq1, seq2, ch.lr * 2.0f / (n_rows_f * world), e == 0, loss_acc, mb_tl_on, mb_tl_step);
            __syncthreads();                                          // new parameters visible; scratch reusable
            MB_TICK(8);
        }
    }
    if (y == 0) {
        for (int i = threadIdx.x; i < NP; i += WS_THREADS) ch.w[i] = S.sw[i];
        if (threadIdx.x == (NP % WS_THREADS) && ch.loss_out) *ch.loss_out = ch.loss_accumulate ? *ch.loss_out + loss_acc : loss_acc;
    }
    tmem_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc_all(*S.tslot);
}

__global__ void __launch_bounds__(WS_THREADS, 1) mb_persist_ws_kernel(const __grid_constant__ MbParams P) {
    extern __shared__ __align__(16) float smem[];
    pdl_launch_dependents();
    int j = 0;
    while (j + 1 < P.n_chains && (int)blockIdx.x >= P.cta_first[j + 1]) ++j;
    const MbChain& ch = P.chains[j];
    const int y = (int)blockIdx.x - P.cta_first[j], gy = P.cta_first[j + 1] - P.cta_first[j];
    if (ch.kind == RCMARL_IN_SA) mb_body_ws<5, 15>(P, ch, j, smem, y, gy);
    else mb_body_ws<5, 10>(P, ch, j, smem, y, gy);
}

#endif  // RCMARL_GRAD_WS

}  // namespace rcmarl
