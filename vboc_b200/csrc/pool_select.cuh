// pool_select.cuh -- device side of the active-learning query on a RESIDENT pool:
// the B most uncertain states of the unlabeled pool by entropy, replacing `np.argpartition(etp, -B)[-B:]` on a
// host copy of the scores (AL/triplependulum_al.py:267-270), and the removal of the queried rows from the pool
// (`np.delete`, :281) as a stable compaction.  All kernels are single passes over the pool at HBM rate
// (the 3-DOF pool is 1.14e7 rows = 45 MB of scores, 273 MB of states).
//
// Top-B = radix select on the order-preserving bit pattern of the (non-negative) scores: three histogram passes
// (11 + 11 + 10 bits) find the key of the B-th largest score, one pass picks everything above it plus as many
// ties as are still needed.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vboc {

constexpr int POOL_THREADS = 256, POOL_PER_THREAD = 16, POOL_BLOCK = POOL_THREADS * POOL_PER_THREAD;

// entropy is >= 0: the IEEE bit pattern orders like the value; NaN / negative scores sort lowest
__device__ __forceinline__ uint32_t pool_key(float v) { return v > 0.f ? __float_as_uint(v) : 0u; }

__global__ void __launch_bounds__(POOL_THREADS) pool_hist_kernel(const float *__restrict__ score, long long P, uint32_t prefix,
                                                                 uint32_t prefix_mask, int shift, int nbits,
                                                                 unsigned int *__restrict__ hist) {
    __shared__ unsigned int sh[2048];
    for (int i = threadIdx.x; i < 2048; i += POOL_THREADS) sh[i] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * POOL_BLOCK;
    const uint32_t bmask = (1u << nbits) - 1u;
#pragma unroll 4
    for (int j = 0; j < POOL_PER_THREAD; ++j) {
        const long long i = base + (long long)j * POOL_THREADS + threadIdx.x;  // coalesced
        if (i < P) {
            const uint32_t key = pool_key(score[i]);
            if ((key & prefix_mask) == prefix) atomicAdd(&sh[(key >> shift) & bmask], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2048; i += POOL_THREADS)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// everything above the k-th key, plus `need` of the ties; sel[0 .. above) / sel[above .. above + need)
__global__ void __launch_bounds__(POOL_THREADS) pool_pick_kernel(const float *__restrict__ score, long long P, uint32_t kth,
                                                                 unsigned int above, unsigned int need,
                                                                 long long *__restrict__ sel, int *__restrict__ flags,
                                                                 unsigned int *__restrict__ counters) {
    const long long base = (long long)blockIdx.x * POOL_BLOCK;
    for (int j = 0; j < POOL_PER_THREAD; ++j) {
        const long long i = base + (long long)j * POOL_THREADS + threadIdx.x;
        if (i >= P) continue;
        const uint32_t key = pool_key(score[i]);
        if (key > kth) {
            const unsigned int pos = atomicAdd(&counters[0], 1u);
            if (pos < above) sel[pos] = i, flags[i] = 1;
        } else if (key == kth) {
            const unsigned int t = atomicAdd(&counters[1], 1u);
            if (t < need) sel[above + t] = i, flags[i] = 1;
        }
    }
}

__global__ void pool_gather_kernel(const float *__restrict__ x, const float *__restrict__ score, const long long *__restrict__ sel,
                                   int k, int n_in, float *__restrict__ gx, float *__restrict__ gscore) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= k) return;
    const long long i = sel[r];
    for (int c = 0; c < n_in; ++c) gx[(size_t)r * n_in + c] = x[(size_t)i * n_in + c];
    gscore[r] = score[i];
}

// rows kept per block of POOL_BLOCK consecutive rows
__global__ void __launch_bounds__(POOL_THREADS) pool_count_kernel(const int *__restrict__ flags, long long P,
                                                                  long long *__restrict__ blk) {
    __shared__ int sh[POOL_THREADS];
    const long long base = (long long)blockIdx.x * POOL_BLOCK + (long long)threadIdx.x * POOL_PER_THREAD;
    int c = 0;
    for (int j = 0; j < POOL_PER_THREAD; ++j)
        if (base + j < P && !flags[base + j]) ++c;
    sh[threadIdx.x] = c;
    __syncthreads();
    for (int o = POOL_THREADS / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) blk[blockIdx.x] = sh[0];
}

// stable compaction: thread t owns POOL_PER_THREAD consecutive rows; exclusive scan of the per-thread kept counts
__global__ void __launch_bounds__(POOL_THREADS) pool_compact_kernel(const float *__restrict__ x, const int *__restrict__ flags,
                                                                    long long P, int n_in, const long long *__restrict__ blkoff,
                                                                    float *__restrict__ xout) {
    __shared__ int sh[POOL_THREADS];
    const long long base = (long long)blockIdx.x * POOL_BLOCK + (long long)threadIdx.x * POOL_PER_THREAD;
    int c = 0;
    for (int j = 0; j < POOL_PER_THREAD; ++j)
        if (base + j < P && !flags[base + j]) ++c;
    sh[threadIdx.x] = c;
    __syncthreads();
    // Hillis-Steele inclusive scan over the 256 per-thread counts
    for (int o = 1; o < POOL_THREADS; o <<= 1) {
        int v = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
        __syncthreads();
        sh[threadIdx.x] += v;
        __syncthreads();
    }
    long long dst = blkoff[blockIdx.x] + sh[threadIdx.x] - c;
    for (int j = 0; j < POOL_PER_THREAD; ++j) {
        const long long i = base + j;
        if (i < P && !flags[i]) {
            for (int q = 0; q < n_in; ++q) xout[(size_t)dst * n_in + q] = x[(size_t)i * n_in + q];
            ++dst;
        }
    }
}

}  // namespace vboc
