// mlp_forward.cuh -- fused inference of the reference's 3-layer MLPs (my_nn.py:4-34: NeuralNetDIR /
// NeuralNetCLS = Linear-ReLU-Linear-ReLU-Linear[-ReLU]) with the input normalisation and the
// label / margin / entropy epilogues of the drivers:
//   mode 1 (VBOC, VBOC/triplependulum_vboc.py:604-620): in = [(q - mean)/std, v/|v|], phi = net(in),
//          label = |v| > phi ? 0 : 1, aux = phi (100 - safety_margin)/100 - |v|   (the margin of
//          VBOC/Safe MPC/parallel/doublependulum_class_fixedveldir.py:240-264)
//   mode 2 (AL, AL/triplependulum_al.py:253-264): in = (x - mean)/std, logits = net(in),
//          aux = scipy.stats.entropy(sigmoid(logits)) (probabilities renormalised to sum 1)
//   mode 0: plain forward of an already normalised input.
// FP32 like the reference (labels flip with the last bits of phi near the boundary, so no reduced
// precision here).  One CTA = 32 rows: layer 1 into shared memory, layer 2 as a register-tiled
// [32 x H] x [H x H] product (thread = output column, 32 row accumulators, W2 pre-transposed so the
// column reads coalesce), layer 3 folded into the epilogue.  The tensor-core (tcgen05) version of the
// H x H product is the planned replacement of `layer2`.
#pragma once
#include <cuda_runtime.h>

namespace vboc {

constexpr int MLP_ROWS = 32, MLP_THREADS = 128, MLP_MAX_IN = 8, MLP_MAX_OUT = 4;

struct MlpParams {
    int batch, n_in, hidden, n_out, mode, final_relu;
    float mean, stdv, margin_scale;  // margin_scale = (100 - safety_margin) / 100
    const float *W1, *b1, *W2T, *b2, *W3, *b3;  // W1 [H][n_in], W2T [H_in][H_out], W3 [n_out][H]
    const float *x;                             // [batch][n_in]
    float *out, *aux;                           // [batch][n_out], [batch]
    int *label;                                 // [batch] (mode 1) or nullptr
};

__global__ void __launch_bounds__(MLP_THREADS) mlp_forward_kernel(const MlpParams P) {
    extern __shared__ float sm[];
    const int H = P.hidden, Hs = H + 1;
    float *a1 = sm;                          // [ROWS][Hs]
    float *a2 = a1 + MLP_ROWS * Hs;          // [ROWS][THREADS + 1]
    float *xin = a2 + MLP_ROWS * (MLP_THREADS + 1);  // [ROWS][MAX_IN]
    float *vn = xin + MLP_ROWS * MLP_MAX_IN;          // [ROWS] velocity norms
    float *o3 = vn + MLP_ROWS;                        // [ROWS][MAX_OUT]
    const int t = threadIdx.x, row0 = blockIdx.x * MLP_ROWS;
    const int n = P.n_in / 2;
    // ---- input normalisation
    if (t < MLP_ROWS) {
        const int r = row0 + t;
        float v[MLP_MAX_IN];
        float nv = 0.f;
        for (int i = 0; i < P.n_in; ++i) v[i] = r < P.batch ? P.x[(size_t)r * P.n_in + i] : 0.f;
        if (P.mode == 1) {
            for (int i = n; i < P.n_in; ++i) nv += v[i] * v[i];
            nv = sqrtf(nv);
            for (int i = 0; i < n; ++i) v[i] = (v[i] - P.mean) / P.stdv;
            if (nv != 0.f)
                for (int i = n; i < P.n_in; ++i) v[i] = v[i] / nv;
        } else if (P.mode == 2) {
            for (int i = 0; i < P.n_in; ++i) v[i] = (v[i] - P.mean) / P.stdv;
        }
        for (int i = 0; i < P.n_in; ++i) xin[t * MLP_MAX_IN + i] = v[i];
        vn[t] = nv;
    }
    if (t < MLP_ROWS * MLP_MAX_OUT) o3[t] = 0.f;
    __syncthreads();
    // ---- layer 1: a1 = relu(x W1' + b1)
    for (int j = t; j < H; j += MLP_THREADS) {
        float w[MLP_MAX_IN];
        for (int i = 0; i < P.n_in; ++i) w[i] = P.W1[j * P.n_in + i];
        const float b = P.b1[j];
        for (int r = 0; r < MLP_ROWS; ++r) {
            float a = b;
            for (int i = 0; i < P.n_in; ++i) a = fmaf(xin[r * MLP_MAX_IN + i], w[i], a);
            a1[r * Hs + j] = fmaxf(a, 0.f);
        }
    }
    __syncthreads();
    // ---- layer 2 (+ layer 3 folded in), THREADS output columns per pass
    for (int j0 = 0; j0 < H; j0 += MLP_THREADS) {
        const int j = j0 + t;
        float acc[MLP_ROWS];
#pragma unroll
        for (int r = 0; r < MLP_ROWS; ++r) acc[r] = 0.f;
        if (j < H) {
            for (int k = 0; k < H; ++k) {
                const float w = P.W2T[(size_t)k * H + j];
#pragma unroll
                for (int r = 0; r < MLP_ROWS; ++r) acc[r] = fmaf(a1[r * Hs + k], w, acc[r]);
            }
            const float b = P.b2[j];
#pragma unroll
            for (int r = 0; r < MLP_ROWS; ++r) a2[r * (MLP_THREADS + 1) + t] = fmaxf(acc[r] + b, 0.f);
        } else {
#pragma unroll
            for (int r = 0; r < MLP_ROWS; ++r) a2[r * (MLP_THREADS + 1) + t] = 0.f;
        }
        __syncthreads();
        // layer 3 partial sums: thread (r, o)
        if (t < MLP_ROWS * P.n_out) {
            const int r = t / P.n_out, o = t - r * P.n_out;
            float a = o3[r * MLP_MAX_OUT + o];
            const int jn = min(MLP_THREADS, H - j0);
            for (int jj = 0; jj < jn; ++jj) a = fmaf(a2[r * (MLP_THREADS + 1) + jj], P.W3[o * H + j0 + jj], a);
            o3[r * MLP_MAX_OUT + o] = a;
        }
        __syncthreads();
    }
    // ---- epilogue
    if (t < MLP_ROWS && row0 + t < P.batch) {
        const int r = row0 + t;
        float o[MLP_MAX_OUT];
        for (int k = 0; k < P.n_out; ++k) {
            float a = o3[t * MLP_MAX_OUT + k] + P.b3[k];
            o[k] = P.final_relu ? fmaxf(a, 0.f) : a;
            P.out[(size_t)r * P.n_out + k] = o[k];
        }
        if (P.mode == 1) {
            if (P.label) P.label[r] = vn[t] > o[0] ? 0 : 1;
            if (P.aux) P.aux[r] = o[0] * P.margin_scale - vn[t];
        } else if (P.mode == 2 && P.aux) {
            float pr[MLP_MAX_OUT], s = 0.f, e = 0.f;
            for (int k = 0; k < P.n_out; ++k) pr[k] = 1.f / (1.f + expf(-o[k])), s += pr[k];
            for (int k = 0; k < P.n_out; ++k) {
                float q = pr[k] / s;
                if (q > 0.f) e -= q * logf(q);
            }
            P.aux[r] = e;
        }
    }
}

inline size_t mlp_smem_bytes(int hidden) {
    return sizeof(float) * (size_t)(MLP_ROWS * (hidden + 1) + MLP_ROWS * (MLP_THREADS + 1) + MLP_ROWS * MLP_MAX_IN +
                                    MLP_ROWS + MLP_ROWS * MLP_MAX_OUT);
}

}  // namespace vboc
