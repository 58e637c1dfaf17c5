// mlp_tc.cuh -- tensor-core version of the fused MLP inference (see mlp_forward.cuh for what is computed).
//
// The H x H layer (the only GEMM-shaped work on the hot path) runs on the 5th-generation tensor cores:
// `tcgen05.mma.cta_group::1.kind::tf32`, M = 128 rows per CTA, N = H (in instructions of N <= 256),
// accumulators in TMEM, operands in shared memory in the canonical K-major no-swizzle layout (8 x 16 B core
// matrices), issued by one thread, completion through `tcgen05.commit` on an mbarrier; the epilogue reads
// the accumulators back with `tcgen05.ld` (thread = row) and folds bias, ReLU, the output layer and the
// label / margin / entropy epilogue in.
// Precision: the reference network is FP32 and labels flip with the last bits of phi near the boundary, so
// the product is done as 3xTF32 (a = a_hi + a_lo with a_hi exactly representable in TF32:
// a_hi b_hi + a_hi b_lo + a_lo b_hi, FP32 accumulation) -- FP32-level accuracy at 3 tensor-core passes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mlp_forward.cuh"

namespace vboc {

constexpr int TC_ROWS = 128, TC_KC = 32, TC_THREADS = 128;

__device__ __forceinline__ uint32_t tc_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE: 8-row core matrices of 16 B rows;
// LBO = byte distance between the two 16 B K-chunks of one MMA, SBO = byte distance between 8-row groups
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version of sm_100
    return d;
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, dense
__device__ __forceinline__ uint32_t tc_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_split(float a, float &hi, float &lo) {
    hi = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u);  // 10 explicit mantissa bits: exact in TF32
    lo = a - hi;
}

// shared memory: [A_hi | A_lo | B_hi | B_lo] operand chunks + small vectors
struct TcLayout {
    int Hp;
    uint32_t a_lbo, b_lbo, a_bytes, b_bytes;
    __host__ __device__ explicit TcLayout(int Hp_) : Hp(Hp_) {
        a_lbo = (TC_ROWS + 1) * 16;  // +1: consecutive K-chunks start in different banks
        b_lbo = (uint32_t)(Hp + 1) * 16;
        a_bytes = a_lbo * (TC_KC / 4);
        b_bytes = b_lbo * (TC_KC / 4);
    }
    __host__ __device__ size_t total() const { return 2 * (size_t)a_bytes + 2 * (size_t)b_bytes + 1024; }
};

// W1p [Hp][n_in], b1p [Hp], W2p [Hp][Hp] (row = output unit, K contiguous), b2p [Hp], W3p [n_out][Hp]: zero padded
__global__ void __launch_bounds__(TC_THREADS, 1) mlp_tc_kernel(const MlpParams P, int Hp, int tmem_cols) {
    extern __shared__ __align__(128) unsigned char smraw[];
    const TcLayout lay(Hp);
    unsigned char *sA_hi = smraw, *sA_lo = sA_hi + lay.a_bytes, *sB_hi = sA_lo + lay.a_bytes, *sB_lo = sB_hi + lay.b_bytes;
    unsigned char *tail = sB_lo + lay.b_bytes;
    unsigned long long *bar = (unsigned long long *)tail;   // mbarrier
    uint32_t *tmem_slot = (uint32_t *)(tail + 16);
    const int t = threadIdx.x, warp = t >> 5, row = blockIdx.x * TC_ROWS + t, n = P.n_in / 2;

    // ---- setup: mbarrier + TMEM allocation (warp 0)
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc_smem(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem(tmem_slot)),
                     "r"((uint32_t)tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // ---- input normalisation (thread = row)
    float xin[MLP_MAX_IN], nv = 0.f;
    for (int i = 0; i < P.n_in; ++i) xin[i] = row < P.batch ? P.x[(size_t)row * P.n_in + i] : 0.f;
    if (P.mode == 1) {
        for (int i = n; i < P.n_in; ++i) nv += xin[i] * xin[i];
        nv = sqrtf(nv);
        for (int i = 0; i < n; ++i) xin[i] = (xin[i] - P.mean) / P.stdv;
        if (nv != 0.f)
            for (int i = n; i < P.n_in; ++i) xin[i] = xin[i] / nv;
    } else if (P.mode == 2) {
        for (int i = 0; i < P.n_in; ++i) xin[i] = (xin[i] - P.mean) / P.stdv;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    uint32_t phase = 0;
    for (int kc0 = 0; kc0 < Hp; kc0 += TC_KC) {
        // ---- A chunk: layer 1 for columns kc0 .. kc0+KC of this thread's row, split hi / lo
#pragma unroll
        for (int c = 0; c < TC_KC / 4; ++c) {
            float hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int k = kc0 + 4 * c + e;
                float a = P.b1[k];
                for (int i = 0; i < P.n_in; ++i) a = fmaf(xin[i], P.W1[k * P.n_in + i], a);
                tc_split(fmaxf(a, 0.f), hi[e], lo[e]);
            }
            *(float4 *)(sA_hi + c * lay.a_lbo + t * 16) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *(float4 *)(sA_lo + c * lay.a_lbo + t * 16) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
        // ---- B chunk: W2p[n][kc0 .. kc0+KC) for all output units n, split hi / lo
        for (int idx = t; idx < Hp * (TC_KC / 4); idx += TC_THREADS) {
            const int c = idx & (TC_KC / 4 - 1), nn = idx / (TC_KC / 4);
            const float4 w = *(const float4 *)(P.W2T + (size_t)nn * Hp + kc0 + 4 * c);  // W2T slot holds W2p here
            float4 h, l;
            tc_split(w.x, h.x, l.x), tc_split(w.y, h.y, l.y), tc_split(w.z, h.z, l.z), tc_split(w.w, h.w, l.w);
            *(float4 *)(sB_hi + c * lay.b_lbo + nn * 16) = h;
            *(float4 *)(sB_lo + c * lay.b_lbo + nn * 16) = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic stores -> tensor-core reads
        __syncthreads();
        // ---- MMA: one thread issues 3 x (K steps) x (N pieces)
        if (t == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int n0 = 0; n0 < Hp;) {
                int nn = Hp - n0;
                if (nn > 256) nn = (Hp > 256 && Hp <= 512 && (Hp / 2) % 16 == 0) ? Hp / 2 : 256;
                const uint32_t idesc = tc_idesc_tf32(TC_ROWS, nn);
#pragma unroll
                for (int ks = 0; ks < TC_KC / 8; ++ks) {
                    const uint32_t aoff = 2 * ks * lay.a_lbo, boff = 2 * ks * lay.b_lbo + n0 * 16;
                    const uint64_t ah = tc_smem_desc(tc_smem(sA_hi) + aoff, lay.a_lbo, 128);
                    const uint64_t al = tc_smem_desc(tc_smem(sA_lo) + aoff, lay.a_lbo, 128);
                    const uint64_t bh = tc_smem_desc(tc_smem(sB_hi) + boff, lay.b_lbo, 128);
                    const uint64_t bl = tc_smem_desc(tc_smem(sB_lo) + boff, lay.b_lbo, 128);
                    const uint32_t first = (kc0 == 0 && ks == 0) ? 0u : 1u;
                    tc_mma_tf32(tmem_base + n0, ah, bh, idesc, first);
                    tc_mma_tf32(tmem_base + n0, ah, bl, idesc, 1u);
                    tc_mma_tf32(tmem_base + n0, al, bh, idesc, 1u);
                }
                n0 += nn;
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                             tc_smem(bar))
                         : "memory");
        }
        // ---- everyone waits until the tensor core has consumed this chunk
        {
            const uint32_t b = tc_smem(bar);
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "TC_WAIT_%=:\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                "@p bra TC_DONE_%=;\n"
                "bra TC_WAIT_%=;\n"
                "TC_DONE_%=:\n"
                "}\n" ::"r"(b),
                "r"(phase)
                : "memory");
            phase ^= 1u;
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: thread = row; accumulators from TMEM, bias + ReLU, output layer, final op
    float o[MLP_MAX_OUT];
    for (int k = 0; k < MLP_MAX_OUT; ++k) o[k] = 0.f;
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < Hp; c0 += 16) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(lane_base + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float v = fmaxf(__uint_as_float(r[j]) + P.b2[c0 + j], 0.f);
            for (int k = 0; k < P.n_out; ++k) o[k] = fmaf(v, P.W3[k * Hp + c0 + j], o[k]);
        }
    }
    if (row < P.batch) {
        for (int k = 0; k < P.n_out; ++k) {
            float a = o[k] + P.b3[k];
            o[k] = P.final_relu ? fmaxf(a, 0.f) : a;
            P.out[(size_t)row * P.n_out + k] = o[k];
        }
        if (P.mode == 1) {
            if (P.label) P.label[row] = nv > o[0] ? 0 : 1;
            if (P.aux) P.aux[row] = o[0] * P.margin_scale - nv;
        } else if (P.mode == 2 && P.aux) {
            float pr[MLP_MAX_OUT], s = 0.f, e = 0.f;
            for (int k = 0; k < P.n_out; ++k) pr[k] = 1.f / (1.f + expf(-o[k])), s += pr[k];
            for (int k = 0; k < P.n_out; ++k) {
                float q = pr[k] / s;
                if (q > 0.f) e -= q * logf(q);
            }
            P.aux[row] = e;
        }
    }
    // ---- release TMEM
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)tmem_cols)
                     : "memory");
}

}  // namespace vboc
