// comm.cuh -- one-shot all-reduce over NVLink peer memory, fused into the reduction / apply kernels
// (data-parallel env shards).  Every rank owns one cudaMalloc'ed exchange buffer that all peers map
// through CUDA IPC:  [ data: 2 slots x world sources x max_floats | flags: one uint32 per peer | CTA arrival counter ].
// Push protocol for step `seq` (identical call sequence on every rank):
//   1. each thread STORES its part of the rank's local sums into slot (seq & 1), source `rank`, of EVERY peer's buffer
//      (fire-and-forget stores over NVLink), fences system-wide, and the CTA bumps the arrival counter; the last CTA to
//      arrive publishes `seq` into flags[rank] of every peer (st.release.sys);
//   2. every CTA waits until its LOCAL flags show `seq` for all peers (local polling, ld.acquire.sys);
//   3. every thread reads the world values of its element from its OWN buffer (no NVLink round trip) and sums them in
//      rank order, so all ranks obtain bitwise identical totals (the replicated parameters never diverge); the SGD
//      update can be applied in the same kernel.
// Double buffering by `seq & 1` is sufficient: a rank can only reach step s+2 (and overwrite slot s & 1 at its peers)
// after every peer has published s+1, which a peer does only after it finished reading step s.
#pragma once
#include <stdint.h>

namespace rcmarl {

constexpr int COMM_MAX_WORLD = 8;

struct CommDev {
    float* data[COMM_MAX_WORLD];       // peer base pointers (data region)
    uint32_t* flags[COMM_MAX_WORLD];   // peer flag arrays [COMM_MAX_WORLD]
    uint32_t* counter;                 // local CTA arrival counter
    uint32_t* error;                   // local error word (spin time-out)
    int64_t max_floats;
    int32_t rank, world;
    uint32_t seq;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_peer(const float* p) {
    float v;
    asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

// Called by all threads of all CTAs of a (small, fully co-resident) grid after they stored their local values
// into comm.data[rank] + slot * max_floats + ...  Returns when every peer's values for this step are readable.
__device__ __forceinline__ void comm_publish_and_wait(const CommDev& c, unsigned total_ctas) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(c.counter, 1u);
        if (t == total_ctas - 1) {
            *c.counter = 0u;
            __threadfence_system();
            for (int p = 0; p < c.world; ++p) st_release_sys(c.flags[p] + c.rank, c.seq);
        }
        const long long t0 = clock64();
        for (int p = 0; p < c.world; ++p) {
            while ((int32_t)(ld_acquire_sys(c.flags[c.rank] + p) - c.seq) < 0) {
                if (clock64() - t0 > 120000000000LL) {   // ~60 s: a peer died; fail loudly instead of hanging the GPU
                    *c.error = 1u;
                    __threadfence_system();
                    __trap();                            // surfaces as a CUDA error at the next synchronisation
                }
            }
        }
    }
    __syncthreads();
}

// element `offset` of source rank `src` in slot (seq & 1) of a buffer
__device__ __forceinline__ int64_t comm_index(const CommDev& c, int src, int64_t offset) {
    return ((int64_t)(c.seq & 1u) * c.world + src) * c.max_floats + offset;
}
// step 1: push this rank's value to every peer (own buffer included)
__device__ __forceinline__ void comm_push(const CommDev& c, int64_t offset, float v) {
    const int64_t idx = comm_index(c, c.rank, offset);
    for (int p = 0; p < c.world; ++p) c.data[p][idx] = v;
}
// step 3: rank-ordered sum of the values all ranks pushed into OUR buffer
__device__ __forceinline__ float comm_total(const CommDev& c, int64_t offset) {
    float v[COMM_MAX_WORLD];
#pragma unroll
    for (int p = 0; p < COMM_MAX_WORLD; ++p)
        v[p] = p < c.world ? ld_peer(c.data[c.rank] + comm_index(c, p, offset)) : 0.f;
    float tot = 0.f;
#pragma unroll
    for (int p = 0; p < COMM_MAX_WORLD; ++p) tot += v[p];   // rank order: identical bits on every rank
    return tot;
}

// host side (comm.cu)
struct CommHost;
CommHost* comm_bound();
bool comm_next(CommDev* out, int64_t need_floats);   // fills *out with the next sequence number; false if unbound

}  // namespace rcmarl
