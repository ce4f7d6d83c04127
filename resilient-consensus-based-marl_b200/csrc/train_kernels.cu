// train_kernels.cu -- update-round kernels of the RPBCAC hot path (sm_100a, fp32 FFMA).
//
//   values_kernel   K1/K3  batched forward values / TD targets / TD errors / actor probabilities
//   grad_kernel     K2/K7/K9  fused forward + backward of the 20-wide MLPs over a row set:
//                   phase 1: one buffer row per lane, weights broadcast from shared memory;
//                   phase 2: the per-row activation / delta vectors are staged in a warp-private
//                            shared-memory tile and every lane owns 4x4 blocks of the weight-gradient
//                            outer products, accumulated in registers across ALL rows of the CTA;
//                   deterministic two-level reduction (CTA partials -> reduce_kernel).
//   team_kernel     K5+K6  neighbour-head estimates, clipped mean, projection numerators
//   consensus, sgd/adam apply, reward mix: the small glue kernels.
//
// Reference semantics: agents/resilient_CAC_agents.py, agents/adversarial_CAC_agents.py,
// training/train_agents.py:86-163 (cited per entry point in include/rcmarl.h).
#include "common.cuh"

namespace rcmarl {

static thread_local int g_last_cuda = 0;
int last_cuda_error_get() { return g_last_cuda; }
void last_cuda_error_set(int e) { g_last_cuda = e; }

int sm_count_cached() {
    static int n = 0;
    if (n == 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
            n = v;
        else
            return 148;  // B200; not cached so that a later call with a device re-queries
    }
    return n;
}

// ============================================================================================
// values
// ============================================================================================
struct ValuesParams {
    rcmarl_rows rows;
    rcmarl_value_job jobs[RCMARL_MAX_JOBS];
};

template <int NA, int DIN>
__device__ __forceinline__ float value_term(const rcmarl_rows& R, const float* sw, int kind, int64_t row) {
    float x[DIN], h1[HID], h2[HID];
    load_x<NA, DIN>(R, kind, row, x);
    features<DIN>(sw, x, h1, h2);
    return head1<DIN>(sw, h2);
}

template <int NA>
__global__ void __launch_bounds__(256) values_kernel(const __grid_constant__ ValuesParams P) {
    extern __shared__ __align__(16) float smem[];
    const rcmarl_value_job& job = P.jobs[blockIdx.y];
    const rcmarl_rows& R = P.rows;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_iter = (R.n_rows + stride - 1) / stride;
    if (job.n_out == NACT) {  // actor.predict: softmax probabilities (single term)
        constexpr int DIN = 2 * NA;
        stage_weights(smem, job.w[0], param_count(DIN, NACT));
        __syncthreads();
        for (int64_t it = 0; it < n_iter; ++it) {
            int64_t m = it * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
            if (m < R.n_rows) {
                int64_t row = row_of(R, m);
                float x[DIN], h1[HID], h2[HID], l[NACT], mx, lse;
                load_x<NA, DIN>(R, job.kind[0], row, x);
                features<DIN>(smem, x, h1, h2);
                head5<DIN>(smem, h2, l);
                if (job.softmax) softmax5(l, mx, lse);
#pragma unroll
                for (int o = 0; o < NACT; ++o) job.out[row * NACT + o] = l[o];
            }
        }
        return;
    }
    for (int t = 0; t < job.n_terms; ++t) {
        const int kind = job.kind[t];
        __syncthreads();
        stage_weights(smem, job.w[t], kind == RCMARL_IN_SA ? param_count(3 * NA, 1) : param_count(2 * NA, 1));
        __syncthreads();
        for (int64_t it = 0; it < n_iter; ++it) {
            int64_t m = it * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
            if (m < R.n_rows) {
                int64_t row = row_of(R, m);
                float v = (kind == RCMARL_IN_SA) ? value_term<NA, 3 * NA>(R, smem, kind, row)
                                                  : value_term<NA, 2 * NA>(R, smem, kind, row);
                float acc;
                if (t == 0)
                    acc = job.add ? job.add_scale * __ldg(job.add + row * job.add_stride + job.add_off) : 0.f;
                else
                    acc = job.out[row];
                job.out[row] = fmaf(job.scale[t], v, acc);
            }
        }
    }
}

// ============================================================================================
// grad
// ============================================================================================
template <int DIN, int NOUT>
struct RowLayout {
    static constexpr int LA1 = round4(DIN + 1);
    static constexpr int OA1 = 0;
    static constexpr int OA2 = LA1;
    static constexpr int OA3 = LA1 + 24;
    static constexpr int OD1 = LA1 + 48;
    static constexpr int OD2 = LA1 + 68;
    static constexpr int OD3 = LA1 + 88;
    static constexpr int LD3 = round4(NOUT);
    static constexpr int RAW = LA1 + 88 + LD3;
    // row stride with an odd number of 16-byte units: conflict-free float4 stores per quarter-warp
    static constexpr int RS = ((RAW / 4) % 2 == 0) ? RAW + 4 : RAW;
    static constexpr int NI1 = LA1 / 4;
    static constexpr int JB3 = LD3 / 4;
    static constexpr int NB1 = NI1 * 5;
    static constexpr int NB2 = 30;
    static constexpr int NB3 = 6 * JB3;
    static constexpr int NBLK = NB1 + NB2 + NB3;
    static constexpr int NPASS = (NBLK + 31) / 32;

    __device__ static __forceinline__ void block_offsets(int b, int& aoff, int& doff) {
        if (b < NB1) {
            aoff = OA1 + 4 * (b / 5); doff = OD1 + 4 * (b % 5);
        } else if (b < NB1 + NB2) {
            b -= NB1; aoff = OA2 + 4 * (b / 5); doff = OD2 + 4 * (b % 5);
        } else if (b < NBLK) {
            b -= NB1 + NB2; aoff = OA3 + 4 * (b / JB3); doff = OD3 + 4 * (b % JB3);
        } else {
            aoff = 0; doff = 0;
        }
    }
    // packed-parameter index of element (ii, jj) of block b, -1 for padding
    __device__ static __forceinline__ int block_param(int b, int ii, int jj) {
        if (b < NB1) {
            int i = 4 * (b / 5) + ii, j = 4 * (b % 5) + jj;
            return i < DIN ? i * HID + j : (i == DIN ? off_b1(DIN) + j : -1);
        } else if (b < NB1 + NB2) {
            b -= NB1;
            int i = 4 * (b / 5) + ii, j = 4 * (b % 5) + jj;
            return i < HID ? off_W2(DIN) + i * HID + j : (i == HID ? off_b2(DIN) + j : -1);
        } else if (b < NBLK) {
            b -= NB1 + NB2;
            int i = 4 * (b / JB3) + ii, o = 4 * (b % JB3) + jj;
            if (o >= NOUT) return -1;
            return i < HID ? off_W3(DIN) + i * NOUT + o : (i == HID ? off_b3(DIN, NOUT) + o : -1);
        }
        return -1;
    }
};

constexpr int GRAD_THREADS = 224;  // 7 warps: two CTAs (2 x ~101 KB of shared memory) per SM at n_agents = 5

struct GradParams {
    rcmarl_rows rows;
    rcmarl_grad_job jobs[RCMARL_MAX_JOBS];
    float* partial;   // [gridDim.y][n_jobs][stride]
    int32_t n_jobs;
    int32_t stride;
};

template <int NA, int DIN, int NOUT>
__device__ __forceinline__ void grad_body(const GradParams& P, const rcmarl_grad_job& job, float* smem) {
    using L = RowLayout<DIN, NOUT>;
    constexpr int NP = param_count(DIN, NOUT);
    rcmarl_rows R = P.rows;
    if (job.time_idx) R.time_idx = job.time_idx;
    float* sw = smem;
    float* tiles = smem + round4(NP);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    float* wt = tiles + warp * (32 * L::RS);
    float* myrow = wt + lane * L::RS;

    stage_weights(sw, job.w, NP);
    __syncthreads();

    int aoff[L::NPASS], doff[L::NPASS];
    float acc[L::NPASS][16];
#pragma unroll
    for (int p = 0; p < L::NPASS; ++p) {
        L::block_offsets(p * 32 + lane, aoff[p], doff[p]);
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[p][e] = 0.f;
    }
    float loss = 0.f;

    const int64_t nchunks = (R.n_rows + 31) >> 5;
    for (int64_t c = (int64_t)blockIdx.y * nwarps + warp; c < nchunks; c += (int64_t)gridDim.y * nwarps) {
        const int64_t m = c * 32 + lane;
        // ---------------- phase 1: one row per lane ----------------
        if (m < R.n_rows) {
            const int64_t row = row_of(R, m);
            float h1[HID], h2[HID];
            {
                float x[DIN];
                load_x<NA, DIN>(R, job.kind, row, x);
                dense20<DIN>(sw, sw + off_b1(DIN), x, h1);
                float4* a1 = reinterpret_cast<float4*>(myrow + L::OA1);
#pragma unroll
                for (int q = 0; q < L::NI1; ++q) {
                    float4 v;
                    v.x = (4 * q + 0 < DIN) ? x[(4 * q + 0 < DIN) ? 4 * q + 0 : 0] : (4 * q + 0 == DIN ? 1.f : 0.f);
                    v.y = (4 * q + 1 < DIN) ? x[(4 * q + 1 < DIN) ? 4 * q + 1 : 0] : (4 * q + 1 == DIN ? 1.f : 0.f);
                    v.z = (4 * q + 2 < DIN) ? x[(4 * q + 2 < DIN) ? 4 * q + 2 : 0] : (4 * q + 2 == DIN ? 1.f : 0.f);
                    v.w = (4 * q + 3 < DIN) ? x[(4 * q + 3 < DIN) ? 4 * q + 3 : 0] : (4 * q + 3 == DIN ? 1.f : 0.f);
                    a1[q] = v;
                }
            }
            dense20<HID>(sw + off_W2(DIN), sw + off_b2(DIN), h1, h2);
            float4* a2 = reinterpret_cast<float4*>(myrow + L::OA2);
            float4* a3 = reinterpret_cast<float4*>(myrow + L::OA3);
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                a2[q] = make_float4(h1[4 * q], h1[4 * q + 1], h1[4 * q + 2], h1[4 * q + 3]);
                a3[q] = make_float4(h2[4 * q], h2[4 * q + 1], h2[4 * q + 2], h2[4 * q + 3]);
            }
            a2[5] = make_float4(1.f, 0.f, 0.f, 0.f);
            a3[5] = make_float4(1.f, 0.f, 0.f, 0.f);

            float d2[HID];
            const float tgt = __ldg(job.target + row * job.target_stride);
            if constexpr (NOUT == 1) {
                // Keras MSE (Appendix A.2): dLoss/dout = 2 (out - y) / B; the 2/B is applied later
                const float e = head1<DIN>(sw, h2) - tgt;
                loss = fmaf(e, e, loss);
                *reinterpret_cast<float4*>(myrow + L::OD3) = make_float4(e, 0.f, 0.f, 0.f);
                const float* W3 = sw + off_W3(DIN);
#pragma unroll
                for (int j = 0; j < HID; ++j) d2[j] = W3[j] * e * lrelu_grad_from_out(h2[j]);
            } else {
                // weighted sparse categorical cross-entropy on the logits (Appendix A.5)
                float p[NACT], mx, lse, g[NACT];
                head5<DIN>(sw, h2, p);
                const int a = (int)__ldg(R.sa + row * (3 * NA) + 3 * job.action_agent + 2);
                float la = 0.f;
#pragma unroll
                for (int o = 0; o < NACT; ++o) la = (o == a) ? p[o] : la;
                softmax5(p, mx, lse);
                loss = fmaf(tgt, (mx + lse) - la, loss);
#pragma unroll
                for (int o = 0; o < NACT; ++o) g[o] = (p[o] - (o == a ? 1.f : 0.f)) * tgt;
                float4* d3 = reinterpret_cast<float4*>(myrow + L::OD3);
                d3[0] = make_float4(g[0], g[1], g[2], g[3]);
                d3[1] = make_float4(g[4], 0.f, 0.f, 0.f);
                const float* W3 = sw + off_W3(DIN);
#pragma unroll
                for (int j = 0; j < HID; ++j) {
                    float s = 0.f;
#pragma unroll
                    for (int o = 0; o < NACT; ++o) s = fmaf(W3[j * NACT + o], g[o], s);
                    d2[j] = s * lrelu_grad_from_out(h2[j]);
                }
            }
            float4* dd2 = reinterpret_cast<float4*>(myrow + L::OD2);
#pragma unroll
            for (int q = 0; q < 5; ++q) dd2[q] = make_float4(d2[4 * q], d2[4 * q + 1], d2[4 * q + 2], d2[4 * q + 3]);
            // d1[i] = (W2[i][:] . d2) * lrelu'(z1[i])
            float4* dd1 = reinterpret_cast<float4*>(myrow + L::OD1);
            const float* W2 = sw + off_W2(DIN);
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                float d1[4];
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) {
                    const int i = 4 * q + ii;
                    const float4* w = reinterpret_cast<const float4*>(W2 + i * HID);
                    float s = 0.f;
#pragma unroll
                    for (int qq = 0; qq < 5; ++qq) {
                        float4 v = w[qq];
                        s = fmaf(v.x, d2[4 * qq + 0], s);
                        s = fmaf(v.y, d2[4 * qq + 1], s);
                        s = fmaf(v.z, d2[4 * qq + 2], s);
                        s = fmaf(v.w, d2[4 * qq + 3], s);
                    }
                    d1[ii] = s * lrelu_grad_from_out(h1[i]);
                }
                dd1[q] = make_float4(d1[0], d1[1], d1[2], d1[3]);
            }
        } else {
            float4* z = reinterpret_cast<float4*>(myrow);
#pragma unroll
            for (int q = 0; q < L::RS / 4; ++q) z[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncwarp();
        // ---------------- phase 2: lane-owned 4x4 blocks of the outer products ----------------
#pragma unroll 4
        for (int r = 0; r < 32; ++r) {
            const float* rp = wt + r * L::RS;
#pragma unroll
            for (int p = 0; p < L::NPASS; ++p) {
                const float4 a = *reinterpret_cast<const float4*>(rp + aoff[p]);
                const float4 d = *reinterpret_cast<const float4*>(rp + doff[p]);
                acc[p][0] = fmaf(a.x, d.x, acc[p][0]);   acc[p][1] = fmaf(a.x, d.y, acc[p][1]);
                acc[p][2] = fmaf(a.x, d.z, acc[p][2]);   acc[p][3] = fmaf(a.x, d.w, acc[p][3]);
                acc[p][4] = fmaf(a.y, d.x, acc[p][4]);   acc[p][5] = fmaf(a.y, d.y, acc[p][5]);
                acc[p][6] = fmaf(a.y, d.z, acc[p][6]);   acc[p][7] = fmaf(a.y, d.w, acc[p][7]);
                acc[p][8] = fmaf(a.z, d.x, acc[p][8]);   acc[p][9] = fmaf(a.z, d.y, acc[p][9]);
                acc[p][10] = fmaf(a.z, d.z, acc[p][10]); acc[p][11] = fmaf(a.z, d.w, acc[p][11]);
                acc[p][12] = fmaf(a.w, d.x, acc[p][12]); acc[p][13] = fmaf(a.w, d.y, acc[p][13]);
                acc[p][14] = fmaf(a.w, d.z, acc[p][14]); acc[p][15] = fmaf(a.w, d.w, acc[p][15]);
            }
        }
        __syncwarp();
    }

    // ---------------- CTA reduction (fixed order => bitwise reproducible) ----------------
    __syncthreads();
    float* red = tiles;  // [nwarps][NPASS*512]
#pragma unroll
    for (int p = 0; p < L::NPASS; ++p) {
        float4* dst = reinterpret_cast<float4*>(red + (warp * L::NPASS + p) * 512 + lane * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            dst[q] = make_float4(acc[p][4 * q], acc[p][4 * q + 1], acc[p][4 * q + 2], acc[p][4 * q + 3]);
    }
    loss = warp_sum(loss);
    float* red_loss = red + nwarps * L::NPASS * 512;
    if (lane == 0) red_loss[warp] = loss;
    __syncthreads();
    float* out = P.partial + ((int64_t)blockIdx.y * P.n_jobs + blockIdx.x) * P.stride;
    for (int q = threadIdx.x; q < L::NPASS * 512; q += blockDim.x) {
        const int p = q >> 9, ln = (q >> 4) & 31, e = q & 15;
        const int idx = L::block_param(p * 32 + ln, e >> 2, e & 3);
        if (idx >= 0) {
            float s = 0.f;
            for (int w = 0; w < nwarps; ++w) s += red[(w * L::NPASS + p) * 512 + ln * 16 + e];
            out[idx] = s;
        }
    }
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < nwarps; ++w) s += red_loss[w];
        out[NP] = s;
    }
}

template <int NA, int LOSS>
__global__ void __launch_bounds__(GRAD_THREADS, (NA <= 5 ? 2 : 1))
grad_kernel(const __grid_constant__ GradParams P) {
    extern __shared__ __align__(16) float smem[];
    const rcmarl_grad_job& job = P.jobs[blockIdx.x];
    if (LOSS == RCMARL_LOSS_CE) {
        grad_body<NA, 2 * NA, NACT>(P, job, smem);
    } else if (job.kind == RCMARL_IN_SA) {
        grad_body<NA, 3 * NA, 1>(P, job, smem);
    } else {
        grad_body<NA, 2 * NA, 1>(P, job, smem);
    }
}

template <int DIN, int NOUT>
constexpr int grad_smem_floats() {
    return round4(param_count(DIN, NOUT)) + (GRAD_THREADS / 32) * 32 * RowLayout<DIN, NOUT>::RS + 64;
}

// sums[j][i] = sum_y partial[y][j][i], y ascending (deterministic)
struct ReduceParams {
    const float* partial;
    float* sums[RCMARL_MAX_JOBS];
    int32_t n[RCMARL_MAX_JOBS];
    int32_t n_jobs, stride, gy;
};
__global__ void __launch_bounds__(256) reduce_kernel(const __grid_constant__ ReduceParams P) {
    const int j = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n[j]) return;
    float s = 0.f;
    for (int y = 0; y < P.gy; ++y) s += P.partial[((int64_t)y * P.n_jobs + j) * P.stride + i];
    P.sums[j][i] = s;
}

// ============================================================================================
// team: estimates + clipped mean + projection numerators
// ============================================================================================
struct TeamParams {
    rcmarl_rows rows;
    rcmarl_team_job jobs[RCMARL_MAX_JOBS];
    float* partial;
    int32_t n_jobs;
    int32_t stride;
};
constexpr int TEAM_N = HID + 2;  // 20 weights + bias numerators + diagnostic loss

template <int NA, int DIN>
__device__ __forceinline__ void team_body(const TeamParams& P, const rcmarl_team_job& job, float* smem) {
    constexpr int NP = param_count(DIN, 1);
    const rcmarl_rows& R = P.rows;
    float* sw = smem;                       // agent's network
    float* heads = smem + round4(NP);       // [n_in][24]: W3 (20), b3, pad
    stage_weights(sw, job.w, NP);
    for (int i = threadIdx.x; i < job.n_in * 24; i += blockDim.x) {
        const int k = i / 24, j = i % 24;
        const float* m = job.msgs + (int64_t)job.in_nodes[k] * job.msg_stride;
        heads[i] = j < HID ? __ldg(m + off_W3(DIN) + j) : (j == HID ? __ldg(m + off_b3(DIN, 1)) : 0.f);
    }
    __syncthreads();
    float acc[TEAM_N];
#pragma unroll
    for (int j = 0; j < TEAM_N; ++j) acc[j] = 0.f;
    const int64_t stride = (int64_t)gridDim.y * blockDim.x;
    const int64_t n_iter = (R.n_rows + stride - 1) / stride;
    for (int64_t it = 0; it < n_iter; ++it) {
        const int64_t m = it * stride + (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
        if (m < R.n_rows) {
            const int64_t row = row_of(R, m);
            float x[DIN], h1[HID], phi[HID];
            load_x<NA, DIN>(R, job.kind, row, x);
            features<DIN>(sw, x, h1, phi);
            float agg;
            if (job.agg_in) {
                agg = __ldg(job.agg_in + row);
            } else {
                float est[RCMARL_MAX_NEIGHBOURS];
#pragma unroll
                for (int k = 0; k < RCMARL_MAX_NEIGHBOURS; ++k) {
                    est[k] = 0.f;
                    if (k < job.n_in) {
                        const float* hk = heads + k * 24;
                        float s = hk[HID];
#pragma unroll
                        for (int j = 0; j < HID; ++j) s = fmaf(phi[j], hk[j], s);
                        est[k] = s;
                    }
                }
                agg = clip_mean_small<RCMARL_MAX_NEIGHBOURS>(est, job.n_in, job.H);
            }
            if (job.agg_out) job.agg_out[row] = agg;
            if (job.sums) {
                const float pred = head1<DIN>(sw, phi);
                float nrm = 1.f;
#pragma unroll
                for (int j = 0; j < HID; ++j) nrm = fmaf(phi[j], phi[j], nrm);
                const float err = agg - pred;
                const float c = err / nrm;
#pragma unroll
                for (int j = 0; j < HID; ++j) acc[j] = fmaf(c, phi[j], acc[j]);
                acc[HID] += c;
                acc[HID + 1] = fmaf(err, c, acc[HID + 1]);
            }
        }
    }
    if (!job.sums) return;
    __syncthreads();
    float* red = smem;  // [nwarps][TEAM_N]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
#pragma unroll
    for (int j = 0; j < TEAM_N; ++j) {
        float s = warp_sum(acc[j]);
        if (lane == 0) red[warp * TEAM_N + j] = s;
    }
    __syncthreads();
    if (threadIdx.x < TEAM_N) {
        float s = 0.f;
        for (int w = 0; w < nwarps; ++w) s += red[w * TEAM_N + threadIdx.x];
        P.partial[((int64_t)blockIdx.y * P.n_jobs + blockIdx.x) * P.stride + threadIdx.x] = s;
    }
}

template <int NA>
__global__ void __launch_bounds__(256) team_kernel(const __grid_constant__ TeamParams P) {
    extern __shared__ __align__(16) float smem[];
    const rcmarl_team_job& job = P.jobs[blockIdx.x];
    if (job.kind == RCMARL_IN_SA) team_body<NA, 3 * NA>(P, job, smem);
    else team_body<NA, 2 * NA>(P, job, smem);
}

// ============================================================================================
// small kernels
// ============================================================================================
struct ConsensusParams { rcmarl_consensus_job jobs[RCMARL_MAX_JOBS]; };
__global__ void __launch_bounds__(256) consensus_hidden_kernel(const __grid_constant__ ConsensusParams P) {
    const rcmarl_consensus_job& job = P.jobs[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= job.n_hidden) return;
    float v[RCMARL_MAX_NEIGHBOURS];
#pragma unroll
    for (int k = 0; k < RCMARL_MAX_NEIGHBOURS; ++k)
        v[k] = k < job.n_in ? __ldg(job.msgs + (int64_t)job.in_nodes[k] * job.msg_stride + i) : 0.f;
    job.dst[i] = clip_mean_small<RCMARL_MAX_NEIGHBOURS>(v, job.n_in, job.H);
}

struct SgdParams { rcmarl_sgd_job jobs[RCMARL_MAX_JOBS]; };
__global__ void __launch_bounds__(256) sgd_kernel(const __grid_constant__ SgdParams P) {
    const rcmarl_sgd_job& job = P.jobs[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < job.n) {
        const float s = job.src[i];
        job.dst[i] = i >= job.first ? s - job.coef * job.sums[i - job.first] : s;
    }
    if (i == 0 && job.loss_out) {
        const float l = job.loss_coef * job.sums[job.n - job.first];
        *job.loss_out = job.loss_accumulate ? *job.loss_out + l : l;
    }
}

struct AdamParams { rcmarl_adam_job jobs[RCMARL_MAX_JOBS]; };
__global__ void __launch_bounds__(256) adam_kernel(const __grid_constant__ AdamParams P) {
    const rcmarl_adam_job& job = P.jobs[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < job.n) {
        const float g = job.grad_scale * job.sums[i];
        const float m = job.beta1 * job.m[i] + (1.f - job.beta1) * g;
        const float v = job.beta2 * job.v[i] + (1.f - job.beta2) * g * g;
        job.m[i] = m;
        job.v[i] = v;
        job.theta[i] = job.theta[i] - job.lr_t * m / (sqrtf(v) + job.eps);
    }
    if (i == 0 && job.loss_out) {
        const float l = job.loss_coef * job.sums[job.n];
        *job.loss_out = job.loss_accumulate ? *job.loss_out + l : l;
    }
}

struct MixParams { int32_t agents[RCMARL_MAX_JOBS]; };
__global__ void __launch_bounds__(256) reward_mix_kernel(const float* __restrict__ r, int64_t n_rows, int n_agents,
                                                         const __grid_constant__ MixParams P, int n_listed,
                                                         float scale, float* __restrict__ out) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const float inv = (float)n_listed;
    float s = 0.f;
    for (int k = 0; k < n_listed; ++k) s = s + __ldg(r + row * n_agents + P.agents[k]) / inv;  // train_agents.py:98
    out[row] = scale * s;
}

// ============================================================================================
// host-side launchers
// ============================================================================================
template <typename K>
static int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) RC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}

static int check_rows(const rcmarl_rows* r) {
    if (!r || !r->sa || !r->ns || !r->r || r->n_rows < 0 || r->n_envs <= 0) return RCMARL_ERR_ARG;
    if (r->n_agents != 5 && r->n_agents != 16) return RCMARL_ERR_ARG;
    if (r->time_idx && (r->n_rows % r->n_envs) != 0) return RCMARL_ERR_ARG;
    return 0;
}

template <int NA>
static int launch_values(const ValuesParams& P, int n_jobs, cudaStream_t st) {
    const size_t smem = sizeof(float) * round4(param_count(3 * NA, 1) > param_count(2 * NA, NACT)
                                                   ? param_count(3 * NA, 1) : param_count(2 * NA, NACT));
    if (set_smem(values_kernel<NA>, smem)) return RCMARL_ERR_CUDA;
    int64_t gx = (P.rows.n_rows + 255) / 256;
    const int64_t cap = (int64_t)sm_count_cached() * 8;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    values_kernel<NA><<<dim3((unsigned)gx, n_jobs), 256, smem, st>>>(P);
    RC_CUDA(cudaGetLastError());
    return 0;
}

template <int NA>
static int launch_grad(GradParams& P, int loss_mode, int gy, cudaStream_t st) {
    constexpr size_t smem_mse = sizeof(float) * (grad_smem_floats<3 * NA, 1>() > grad_smem_floats<2 * NA, 1>()
                                                     ? grad_smem_floats<3 * NA, 1>() : grad_smem_floats<2 * NA, 1>());
    constexpr size_t smem_ce = sizeof(float) * grad_smem_floats<2 * NA, NACT>();
    if (loss_mode == RCMARL_LOSS_CE) {
        if (set_smem(grad_kernel<NA, RCMARL_LOSS_CE>, smem_ce)) return RCMARL_ERR_CUDA;
        grad_kernel<NA, RCMARL_LOSS_CE><<<dim3(P.n_jobs, gy), GRAD_THREADS, smem_ce, st>>>(P);
    } else {
        if (set_smem(grad_kernel<NA, RCMARL_LOSS_MSE>, smem_mse)) return RCMARL_ERR_CUDA;
        grad_kernel<NA, RCMARL_LOSS_MSE><<<dim3(P.n_jobs, gy), GRAD_THREADS, smem_mse, st>>>(P);
    }
    RC_CUDA(cudaGetLastError());
    return 0;
}

template <int NA>
static int launch_team(const TeamParams& P, int gy, cudaStream_t st) {
    const size_t smem = sizeof(float) * (round4(param_count(3 * NA, 1)) + RCMARL_MAX_NEIGHBOURS * 24 + 8 * TEAM_N);
    if (set_smem(team_kernel<NA>, smem)) return RCMARL_ERR_CUDA;
    team_kernel<NA><<<dim3(P.n_jobs, gy), 256, smem, st>>>(P);
    RC_CUDA(cudaGetLastError());
    return 0;
}

static int grid_y_for(int64_t work_items, int n_jobs, int ctas_per_sm) {
    int64_t cap = ((int64_t)sm_count_cached() * ctas_per_sm + n_jobs - 1) / n_jobs;
    if (cap < 1) cap = 1;
    int64_t gy = work_items < cap ? work_items : cap;
    return (int)(gy < 1 ? 1 : gy);
}

}  // namespace rcmarl

using namespace rcmarl;

extern "C" {

int64_t rcmarl_workspace_bytes(int n_jobs, int max_params) {
    if (n_jobs < 1) n_jobs = 1;
    const int64_t ctas = (int64_t)sm_count_cached() * 2 + 2 * (int64_t)n_jobs;
    return ctas * (int64_t)(max_params + 1) * (int64_t)sizeof(float);
}

int rcmarl_values(const rcmarl_rows* rows, const rcmarl_value_job* jobs, int n_jobs, void* stream) {
    if (int e = check_rows(rows)) return e;
    if (!jobs || n_jobs < 1 || n_jobs > RCMARL_MAX_JOBS) return RCMARL_ERR_ARG;
    if (rows->n_rows == 0) return RCMARL_OK;
    ValuesParams P;
    P.rows = *rows;
    for (int j = 0; j < n_jobs; ++j) {
        const rcmarl_value_job& q = jobs[j];
        if (!q.out || q.n_terms < 1 || q.n_terms > RCMARL_MAX_TERMS) return RCMARL_ERR_ARG;
        if (q.n_out != 1 && !(q.n_out == NACT && q.n_terms == 1 && q.kind[0] != RCMARL_IN_SA)) return RCMARL_ERR_ARG;
        for (int t = 0; t < q.n_terms; ++t)
            if (!q.w[t] || q.kind[t] < 0 || q.kind[t] > 2) return RCMARL_ERR_ARG;
        P.jobs[j] = q;
    }
    cudaStream_t st = (cudaStream_t)stream;
    return rows->n_agents == 5 ? launch_values<5>(P, n_jobs, st) : launch_values<16>(P, n_jobs, st);
}

int rcmarl_grad(const rcmarl_rows* rows, const rcmarl_grad_job* jobs, int n_jobs, int loss_mode, void* ws,
                int64_t ws_bytes, void* stream) {
    if (int e = check_rows(rows)) return e;
    if (!jobs || n_jobs < 1 || n_jobs > RCMARL_MAX_JOBS || !ws) return RCMARL_ERR_ARG;
    if (loss_mode != RCMARL_LOSS_MSE && loss_mode != RCMARL_LOSS_CE) return RCMARL_ERR_ARG;
    const int NA = rows->n_agents;
    GradParams P;
    ReduceParams Q;
    P.rows = *rows;
    int maxn = 0;
    for (int j = 0; j < n_jobs; ++j) {
        const rcmarl_grad_job& q = jobs[j];
        if (!q.w || !q.target || !q.sums || q.kind < 0 || q.kind > 2 || q.target_stride < 1) return RCMARL_ERR_ARG;
        if (q.time_idx && !rows->time_idx) return RCMARL_ERR_ARG;   /* overrides need the gathered row mode */
        if (loss_mode == RCMARL_LOSS_CE && (q.kind != RCMARL_IN_S || q.action_agent < 0 || q.action_agent >= NA))
            return RCMARL_ERR_ARG;
        P.jobs[j] = q;
        const int n = loss_mode == RCMARL_LOSS_CE ? param_count(2 * NA, NACT)
                                                   : param_count(q.kind == RCMARL_IN_SA ? 3 * NA : 2 * NA, 1);
        Q.sums[j] = q.sums;
        Q.n[j] = n + 1;
        if (n + 1 > maxn) maxn = n + 1;
    }
    const int64_t nchunks = (rows->n_rows + 31) / 32;
    const int gy = grid_y_for((nchunks + GRAD_THREADS / 32 - 1) / (GRAD_THREADS / 32), n_jobs, NA <= 5 ? 2 : 1);
    if ((int64_t)gy * n_jobs * maxn * (int64_t)sizeof(float) > ws_bytes) return RCMARL_ERR_WORKSPACE;
    P.partial = (float*)ws;
    P.n_jobs = n_jobs;
    P.stride = maxn;
    cudaStream_t st = (cudaStream_t)stream;
    int e = NA == 5 ? launch_grad<5>(P, loss_mode, gy, st) : launch_grad<16>(P, loss_mode, gy, st);
    if (e) return e;
    Q.partial = (const float*)ws;
    Q.n_jobs = n_jobs;
    Q.stride = maxn;
    Q.gy = gy;
    reduce_kernel<<<dim3((maxn + 255) / 256, n_jobs), 256, 0, st>>>(Q);
    RC_CUDA(cudaGetLastError());
    return RCMARL_OK;
}

int rcmarl_team(const rcmarl_rows* rows, const rcmarl_team_job* jobs, int n_jobs, void* ws, int64_t ws_bytes,
                void* stream) {
    if (int e = check_rows(rows)) return e;
    if (!jobs || n_jobs < 1 || n_jobs > RCMARL_MAX_JOBS) return RCMARL_ERR_ARG;
    const int NA = rows->n_agents;
    TeamParams P;
    ReduceParams Q;
    P.rows = *rows;
    bool any_sums = false;
    for (int j = 0; j < n_jobs; ++j) {
        const rcmarl_team_job& q = jobs[j];
        if (!q.w || q.kind < 0 || q.kind > 1) return RCMARL_ERR_ARG;
        if (!q.agg_in) {
            if (!q.msgs || q.n_in < 1 || q.n_in > RCMARL_MAX_NEIGHBOURS || q.H < 0 || q.H >= q.n_in) return RCMARL_ERR_ARG;
        }
        if (!q.sums && !q.agg_out) return RCMARL_ERR_ARG;
        any_sums |= (q.sums != nullptr);
        P.jobs[j] = q;
        Q.sums[j] = q.sums;
        Q.n[j] = q.sums ? TEAM_N : 0;
    }
    const int gy = grid_y_for((rows->n_rows + 255) / 256, n_jobs, 4);
    if (any_sums) {
        if (!ws || (int64_t)gy * n_jobs * TEAM_N * (int64_t)sizeof(float) > ws_bytes) return RCMARL_ERR_WORKSPACE;
    }
    P.partial = (float*)ws;
    P.n_jobs = n_jobs;
    P.stride = TEAM_N;
    cudaStream_t st = (cudaStream_t)stream;
    int e = NA == 5 ? launch_team<5>(P, gy, st) : launch_team<16>(P, gy, st);
    if (e) return e;
    if (any_sums) {
        Q.partial = (const float*)ws;
        Q.n_jobs = n_jobs;
        Q.stride = TEAM_N;
        Q.gy = gy;
        reduce_kernel<<<dim3(1, n_jobs), 256, 0, st>>>(Q);
        RC_CUDA(cudaGetLastError());
    }
    return RCMARL_OK;
}

int rcmarl_consensus_hidden(const rcmarl_consensus_job* jobs, int n_jobs, void* stream) {
    if (!jobs || n_jobs < 1 || n_jobs > RCMARL_MAX_JOBS) return RCMARL_ERR_ARG;
    ConsensusParams P;
    int maxn = 0;
    for (int j = 0; j < n_jobs; ++j) {
        const rcmarl_consensus_job& q = jobs[j];
        if (!q.dst || !q.msgs || q.n_in < 1 || q.n_in > RCMARL_MAX_NEIGHBOURS || q.H < 0 || q.H >= q.n_in || q.n_hidden < 1)
            return RCMARL_ERR_ARG;
        P.jobs[j] = q;
        if (q.n_hidden > maxn) maxn = q.n_hidden;
    }
    consensus_hidden_kernel<<<dim3((maxn + 255) / 256, n_jobs), 256, 0, (cudaStream_t)stream>>>(P);
    RC_CUDA(cudaGetLastError());
    return RCMARL_OK;
}

int rcmarl_sgd_apply(const rcmarl_sgd_job* jobs, int n_jobs, void* stream) {
    if (!jobs || n_jobs < 1 || n_jobs > RCMARL_MAX_JOBS) return RCMARL_ERR_ARG;
    SgdParams P;
    int maxn = 0;
    for (int j = 0; j < n_jobs; ++j) {
        if (!jobs[j].dst || !jobs[j].src || !jobs[j].sums || jobs[j].n < 1) return RCMARL_ERR_ARG;
        P.jobs[j] = jobs[j];
        if (jobs[j].n > maxn) maxn = jobs[j].n;
    }
    sgd_kernel<<<dim3((maxn + 255) / 256, n_jobs), 256, 0, (cudaStream_t)stream>>>(P);
    RC_CUDA(cudaGetLastError());
    return RCMARL_OK;
}

int rcmarl_adam_apply(const rcmarl_adam_job* jobs, int n_jobs, void* stream) {
    if (!jobs || n_jobs < 1 || n_jobs > RCMARL_MAX_JOBS) return RCMARL_ERR_ARG;
    AdamParams P;
    int maxn = 0;
    for (int j = 0; j < n_jobs; ++j) {
        if (!jobs[j].theta || !jobs[j].m || !jobs[j].v || !jobs[j].sums || jobs[j].n < 1) return RCMARL_ERR_ARG;
        P.jobs[j] = jobs[j];
        if (jobs[j].n > maxn) maxn = jobs[j].n;
    }
    adam_kernel<<<dim3((maxn + 255) / 256, n_jobs), 256, 0, (cudaStream_t)stream>>>(P);
    RC_CUDA(cudaGetLastError());
    return RCMARL_OK;
}

int rcmarl_reward_mix(const float* r, int64_t n_rows, int n_agents, const int32_t* agents, int n_listed, float scale,
                      float* out, void* stream) {
    if (!r || !out || !agents || n_rows < 0 || n_listed < 1 || n_listed > RCMARL_MAX_JOBS || n_agents < 1)
        return RCMARL_ERR_ARG;
    if (n_rows == 0) return RCMARL_OK;
    MixParams P;
    for (int k = 0; k < n_listed; ++k) {
        if (agents[k] < 0 || agents[k] >= n_agents) return RCMARL_ERR_ARG;
        P.agents[k] = agents[k];
    }
    reward_mix_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(r, n_rows, n_agents, P,
                                                                                         n_listed, scale, out);
    RC_CUDA(cudaGetLastError());
    return RCMARL_OK;
}

}  // extern "C"
