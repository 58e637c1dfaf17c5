// mlp_pipe.cuh -- the tensor-core MLP inference as a software pipeline (round 2; mlp_tc.cuh is the serial version
// it replaces and stays as the cross-check, VBOC_MLP_SERIAL=1).
//
// What is computed: mlp_forward.cuh (my_nn.py:4-34 with the drivers' normalisation and label / margin / entropy
// epilogues).  The H x H layer runs as 3xTF32 on tcgen05 (a = a_hi + a_lo: a_hi b_hi + a_hi b_lo + a_lo b_hi, FP32
// accumulation in TMEM), M = 128 rows per tile, K in chunks of 8 (one MMA K-step).
//
// Pipeline, one persistent CTA per SM looping over row tiles:
//   * W2 is split into hi / lo ONCE (vboc_mlp_create) and stored in global memory as the exact shared-memory image
//     of every K-chunk (canonical K-major no-swizzle core matrices), so a chunk of the B operand is one
//     `cp.async.bulk` (TMA, SASS UBLKCP) into a 4-slot ring, two chunks ahead of the MMA that reads it;
//   * the A operand (layer 1 of the tile, hi / lo) is produced by the 128 threads (thread = row) into a 2-slot
//     ring from W1 / b1 held in shared memory, one chunk ahead: while the tensor core works on chunk g the
//     threads build chunk g + 1;
//   * one elected thread issues the `tcgen05.mma` of a chunk and commits them to the mbarrier that frees the
//     chunk's A and B slots;
//   * the epilogue (tcgen05.ld, bias, ReLU, output layer, label / margin / entropy) runs thread = row.
// All of TMEM (512 columns) holds the 128 x Hp accumulator of ONE tile when Hp = 512, so the epilogue of a tile
// cannot overlap the MMAs of the next one; it is ~10 % of a tile.
#pragma once
#include "mlp_tc.cuh"

namespace vboc {

constexpr int TP_KC = 8, TP_BSLOTS = 4, TP_ASLOTS = 2;

struct TpLayout {
    int Hp, n_in, n_out;
    uint32_t a_lbo, b_lbo, a_bytes, b_bytes;  // bytes of ONE (hi or lo) chunk image
    __host__ __device__ TpLayout(int Hp_, int n_in_, int n_out_) : Hp(Hp_), n_in(n_in_), n_out(n_out_) {
        a_lbo = (TC_ROWS + 1) * 16;
        b_lbo = (uint32_t)(Hp + 1) * 16;
        a_bytes = a_lbo * (TP_KC / 4);
        b_bytes = b_lbo * (TP_KC / 4);
    }
    __host__ __device__ size_t off_a(int slot) const { return (size_t)slot * 2 * a_bytes; }
    __host__ __device__ size_t off_b(int slot) const { return (size_t)TP_ASLOTS * 2 * a_bytes + (size_t)slot * 2 * b_bytes; }
    __host__ __device__ size_t off_w() const { return off_b(TP_BSLOTS); }  // W1 [Hp][n_in], b1, b2 [Hp], W3 [n_out][Hp]
    __host__ __device__ size_t w_floats() const { return (size_t)Hp * n_in + 2 * (size_t)Hp + (size_t)n_out * Hp; }
    __host__ __device__ size_t off_bar() const { return (off_w() + w_floats() * 4 + 15) & ~(size_t)15; }
    __host__ __device__ size_t total() const { return off_bar() + 128; }
};

// bounded wait: a pipeline bug must end in a trap (an error the host sees), never in a kernel that spins forever
__device__ __forceinline__ void tp_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (int spins = 0; !done; ++spins) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && spins > (1 << 24)) __trap();
    }
}

// W2img: [Hp / 8 chunks][hi | lo][b_bytes]: the shared-memory images of the B operand chunks
// NIN / NOUT are compile-time (2n in {2, 4, 6}, 1 or 2 outputs: every network of the reference) so that the row's
// inputs and outputs live in registers and the layer-1 / output-layer loops unroll.
template <int NIN, int NOUT>
__global__ void __launch_bounds__(TC_THREADS, 1) mlp_pipe_kernel(const MlpParams P, int Hp, int tmem_cols,
                                                                 const unsigned char *__restrict__ W2img, long long batch) {
    extern __shared__ __align__(128) unsigned char smraw[];
    const TpLayout lay(Hp, NIN, NOUT);
    float *sW1 = (float *)(smraw + lay.off_w()), *sb1 = sW1 + (size_t)Hp * NIN, *sb2 = sb1 + Hp, *sW3 = sb2 + Hp;
    unsigned long long *bars = (unsigned long long *)(smraw + lay.off_bar());
    // bars[0..3] B slot full, bars[4..5] MMA of the A slot done, bars[8] (as uint32) TMEM base
    uint32_t *tmem_slot = (uint32_t *)(bars + 8);
    const int t = threadIdx.x, warp = t >> 5;
    constexpr int n = NIN / 2;
    const int NCH = Hp / TP_KC;
    const long long tiles = (batch + TC_ROWS - 1) / TC_ROWS;
    const uint32_t bfull0 = tc_smem(bars), mdone0 = tc_smem(bars + 4);

    if (t == 0) {
        for (int i = 0; i < TP_BSLOTS; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bfull0 + 8u * i));
        for (int i = 0; i < TP_ASLOTS; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mdone0 + 8u * i));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem(tmem_slot)),
                     "r"((uint32_t)tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // small weights into shared memory (zero padded copies: W1p, b1p, b2p, W3p)
    for (int i = t; i < Hp * NIN; i += TC_THREADS) sW1[i] = P.W1[i];
    for (int i = t; i < Hp; i += TC_THREADS) sb1[i] = P.b1[i], sb2[i] = P.b2[i];
    for (int i = t; i < NOUT * Hp; i += TC_THREADS) sW3[i] = P.W3[i];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    const long long my_tiles = blockIdx.x < tiles ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long total_chunks = my_tiles * NCH;
    // B ring: chunk g of this CTA's sequence is K-chunk g % NCH (the same W2 images for every tile)
    auto fetch_b = [&](long long g) {
        const int q = (int)(g & (TP_BSLOTS - 1));
        const uint32_t bar = bfull0 + 8u * q, bytes = 2u * lay.b_bytes;
        const unsigned char *src = W2img + (size_t)(g % NCH) * bytes;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         tc_smem(smraw + lay.off_b(q))),
                     "l"(src), "r"(bytes), "r"(bar)
                     : "memory");
    };
    if (t == 0) {
        if (total_chunks > 0) fetch_b(0);
        if (total_chunks > 1) fetch_b(1);
    }

    long long g = 0;  // chunk counter of this CTA
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long row = tile * TC_ROWS + t;
        // ---- input normalisation (thread = row)
        float xin[NIN], nv = 0.f;
#pragma unroll
        for (int i = 0; i < NIN; ++i) xin[i] = row < batch ? P.x[(size_t)row * NIN + i] : 0.f;
        if (P.mode == 1) {
#pragma unroll
            for (int i = n; i < NIN; ++i) nv += xin[i] * xin[i];
            nv = sqrtf(nv);
#pragma unroll
            for (int i = 0; i < n; ++i) xin[i] = (xin[i] - P.mean) / P.stdv;
            if (nv != 0.f) {
#pragma unroll
                for (int i = n; i < NIN; ++i) xin[i] = xin[i] / nv;
            }
        } else if (P.mode == 2) {
#pragma unroll
            for (int i = 0; i < NIN; ++i) xin[i] = (xin[i] - P.mean) / P.stdv;
        }
        for (int ch = 0; ch < NCH; ++ch, ++g) {
            const int as = (int)(g & (TP_ASLOTS - 1)), q = (int)(g & (TP_BSLOTS - 1));
            // the A slot (and the B slot of chunk g - 2) is free once the MMAs of chunk g - 2 have completed
            if (g >= TP_ASLOTS) tp_wait(mdone0 + 8u * as, (uint32_t)(((g >> 1) - 1) & 1));
            if (t == 0 && g + 2 < total_chunks) fetch_b(g + 2);  // into the B slot chunk g - 2 used
            // ---- A chunk: layer 1 for hidden units ch*8 .. ch*8+7 of this thread's row, split hi / lo
            unsigned char *sA_hi = smraw + lay.off_a(as), *sA_lo = sA_hi + lay.a_bytes;
#pragma unroll
            for (int c = 0; c < TP_KC / 4; ++c) {
                float hi[4], lo[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int k = ch * TP_KC + 4 * c + e;
                    float a = sb1[k];
#pragma unroll
                    for (int i = 0; i < NIN; ++i) a = fmaf(xin[i], sW1[k * NIN + i], a);
                    tc_split(fmaxf(a, 0.f), hi[e], lo[e]);
                }
                *(float4 *)(sA_hi + c * lay.a_lbo + t * 16) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *(float4 *)(sA_lo + c * lay.a_lbo + t * 16) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic stores -> tensor-core reads
            __syncthreads();
            if (t == 0) {
                tp_wait(bfull0 + 8u * q, (uint32_t)((g >> 2) & 1));  // the chunk's W2 image has landed
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned char *sB_hi = smraw + lay.off_b(q), *sB_lo = sB_hi + lay.b_bytes;
                for (int n0 = 0; n0 < Hp;) {
                    int nn = Hp - n0;
                    if (nn > 256) nn = (Hp > 256 && Hp <= 512 && (Hp / 2) % 16 == 0) ? Hp / 2 : 256;
                    const uint32_t idesc = tc_idesc_tf32(TC_ROWS, nn);
                    const uint64_t ah = tc_smem_desc(tc_smem(sA_hi), lay.a_lbo, 128);
                    const uint64_t al = tc_smem_desc(tc_smem(sA_lo), lay.a_lbo, 128);
                    const uint64_t bh = tc_smem_desc(tc_smem(sB_hi) + n0 * 16, lay.b_lbo, 128);
                    const uint64_t bl = tc_smem_desc(tc_smem(sB_lo) + n0 * 16, lay.b_lbo, 128);
                    tc_mma_tf32(tmem_base + n0, ah, bh, idesc, ch == 0 ? 0u : 1u);
                    tc_mma_tf32(tmem_base + n0, ah, bl, idesc, 1u);
                    tc_mma_tf32(tmem_base + n0, al, bh, idesc, 1u);
                    n0 += nn;
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                 mdone0 + 8u * as)
                             : "memory");
            }
        }
        // ---- all MMAs of the tile done: the last two commits cover every earlier one (in-order completion)
        tp_wait(mdone0 + 8u * (uint32_t)((g - 1) & 1), (uint32_t)(((g - 1) >> 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

        // ---- epilogue: thread = row; accumulators from TMEM, bias + ReLU, output layer, final op
        float o[NOUT];
#pragma unroll
        for (int k = 0; k < NOUT; ++k) o[k] = 0.f;
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
                const float v = fmaxf(__uint_as_float(r[j]) + sb2[c0 + j], 0.f);
#pragma unroll
                for (int k = 0; k < NOUT; ++k) o[k] = fmaf(v, sW3[k * Hp + c0 + j], o[k]);
            }
        }
        if (row < batch) {
#pragma unroll
            for (int k = 0; k < NOUT; ++k) {
                float a = o[k] + P.b3[k];
                o[k] = P.final_relu ? fmaxf(a, 0.f) : a;
                if (P.out) P.out[(size_t)row * NOUT + k] = o[k];
            }
            if (P.mode == 1) {
                if (P.label) P.label[row] = nv > o[0] ? 0 : 1;
                if (P.aux) P.aux[row] = o[0] * P.margin_scale - nv;
            } else if (P.mode == 2 && P.aux) {
                float pr[NOUT], s = 0.f, e = 0.f;
#pragma unroll
                for (int k = 0; k < NOUT; ++k) pr[k] = 1.f / (1.f + expf(-o[k])), s += pr[k];
#pragma unroll
                for (int k = 0; k < NOUT; ++k) {
                    float q = pr[k] / s;
                    if (q > 0.f) e -= q * logf(q);
                }
                P.aux[row] = e;
            }
        }
        // the next tile's first MMA overwrites the accumulators: every thread's TMEM reads come first
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    // ---- release TMEM
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)tmem_cols)
                     : "memory");
}

}  // namespace vboc
