// nn_margin.h -- the learned viability margin as a constraint function INSIDE the OCP.
//
// Stands in for `nn_decisionfunction` of the reference's Safe-MPC classes
// (VBOC/Safe MPC/hard_terminal_constraints/doublependulum_class_fixedveldir.py:232-258 and
//  VBOC/Safe MPC/parallel/doublependulum_class_fixedveldir.py:240-264), which CasADi expands symbolically and acados
// differentiates:
//     vn  = max(|v|, 1e-3)
//     in  = [(q - mean) / std, v / vn]
//     out = W3 relu(W2 relu(W1 in + b1) + b2) + b3          (no ReLU on the output in the constraint function)
//     h(x) = out * (100 - safety_margin) / 100 - vn  >= 0    (the state is inside the learned viability kernel)
// evaluated in FP64 on the FP32-trained weights, with its gradient dh/dx by reverse mode through the two hidden
// layers and the input normalisation.  One warp evaluates it: lanes stride over the hidden units, the activations
// live in the warp's global workspace (2 x H doubles, L1 resident), the 2n inputs / outputs go through shared memory.
#pragma once
#include "warp_spmd.h"

namespace vboc {

constexpr int NN_HMAX = 512;

// device-resident network (double copies of the weights; W2 in both layouts so that both passes read coalesced)
struct NnNet {
    int n_in, hidden;
    const double *W1;   // [H][n_in]
    const double *b1;   // [H]
    const double *W2;   // [H_out][H_in]
    const double *W2T;  // [H_in][H_out]
    const double *b2;   // [H]
    const double *W3;   // [H]   (one output)
    double b3;
    double mean, stdv, scale;  // scale = (100 - safety_margin) / 100
    // vel_norm = |x[vstart:]|.  n_dof for the double-pendulum classes; the triple-pendulum classes write `x[2:]` as well
    // (VBOC/Safe MPC/triplependulum_class_vboc.py:217, 282), which includes theta_3: vstart = 2 mirrors them
    int vstart;
};

// Guess network of the AL drivers (AL/triplependulum_class_al.py:171-201, `compute_problem_nnguess`): a 2n-H-H-(N 2n)
// MLP (my_nn.py NeuralNetCLS) that predicts the whole state trajectory from the initial state; forward only.
struct GuessNet {
    int hidden, n_out;
    const double *W1;   // [H][n_in]
    const double *b1;   // [H]
    const double *W2T;  // [H_in][H_out]
    const double *b2;   // [H]
    const double *W3T;  // [H][n_out]
    const double *b3;   // [n_out]
    double mean, stdv;
};

// out[j] = (W3 relu(W2 relu(W1 (x0 - mean)/std + b1) + b2) + b3)[j] * std + mean, j < n_out, written through `put(j, v)`.
// x0 in shared memory, a1 / a2 (H doubles each) in global memory.
template <int NQ, class Put>
VB_DEV void guess_forward(const GuessNet &net, const double *x0, double *a1, double *a2, Put put) {
    constexpr int NX = 2 * NQ;
    const int H = net.hidden;
    FOR_LANES
    for (int k = lane; k < H; k += 32) {
        double a = net.b1[k];
#pragma unroll
        for (int i = 0; i < NX; ++i) a += net.W1[k * NX + i] * ((x0[i] - net.mean) / net.stdv);
        a1[k] = a > 0.0 ? a : 0.0;
    }
    END_LANES
    FOR_LANES
    for (int j = lane; j < H; j += 32) {
        double a = net.b2[j];
        for (int k = 0; k < H; ++k) a += net.W2T[(size_t)k * H + j] * a1[k];
        a2[j] = a > 0.0 ? a : 0.0;
    }
    END_LANES
    FOR_LANES
    for (int j = lane; j < net.n_out; j += 32) {
        double a = net.b3[j];
        for (int k = 0; k < H; ++k) a += net.W3T[(size_t)k * net.n_out + j] * a2[k];
        put(j, a * net.stdv + net.mean);
    }
    END_LANES
}

// h(x) and, if grad != nullptr, dh/dx (NX doubles).  x, grad and the scratch `io` (>= 2 NX + 2 doubles) are in shared
// memory, a1 / a2 (H doubles each) in global memory.  Warp-uniform result.
template <int NQ>
VB_DEV double nn_margin(const NnNet &net, const double *x, double *grad, double *a1, double *a2, double *io) {
    constexpr int NX = 2 * NQ;
    const int H = net.hidden;
    double vn2 = 0.0;
#pragma unroll
    for (int i = 0; i < NX; ++i)
        if (i >= net.vstart) vn2 += x[i] * x[i];
    const double vraw = sqrt(vn2);
    const bool clip = !(vraw > 1e-3);
    const double vn = clip ? 1e-3 : vraw;
    double *in = io, *gin = io + NX;  // normalised input, gradient wrt it
    FOR_LANES
    if (lane < NQ) in[lane] = (x[lane] - net.mean) / net.stdv;
    else if (lane < NX) in[lane] = x[lane] / vn;
    END_LANES
    // layer 1 (+ ReLU)
    FOR_LANES
    for (int k = lane; k < H; k += 32) {
        double a = net.b1[k];
#pragma unroll
        for (int i = 0; i < NX; ++i) a += net.W1[k * NX + i] * in[i];
        a1[k] = a > 0.0 ? a : 0.0;
    }
    END_LANES
    // layer 2 (+ ReLU) and the output, lanes over the output units (W2T: coalesced)
    LV(double, acc);
    FOR_LANES
    double o = 0.0;
    for (int j = lane; j < H; j += 32) {
        double a = net.b2[j];
        for (int k = 0; k < H; ++k) a += net.W2T[(size_t)k * H + j] * a1[k];
        a = a > 0.0 ? a : 0.0;
        a2[j] = a;
        o += net.W3[j] * a;
    }
    L(acc) = o;
    END_LANES
    const double out = WARP_SUM(acc) + net.b3;
    const double h = out * net.scale - vn;
    if (grad) {
        // reverse mode: d2_j = relu'(a2_j) W3_j ; d1_k = relu'(a1_k) sum_j W2[j][k] d2_j ; gin_i = sum_k W1[k][i] d1_k
        FOR_LANES
        for (int j = lane; j < H; j += 32) a2[j] = a2[j] > 0.0 ? net.W3[j] : 0.0;
        END_LANES
        FOR_LANES
        for (int k = lane; k < H; k += 32) {
            double d = 0.0;
            if (a1[k] > 0.0)
                for (int j = 0; j < H; ++j) d += net.W2[(size_t)j * H + k] * a2[j];
            a1[k] = d;
        }
        END_LANES
        LV(double, g0);
#pragma unroll 1
        for (int i = 0; i < NX; ++i) {
            FOR_LANES
            double d = 0.0;
            for (int k = lane; k < H; k += 32) d += net.W1[k * NX + i] * a1[k];
            L(g0) = d;
            END_LANES
            const double gi = WARP_SUM(g0);
            FOR_LANES
            if (lane == 0) gin[i] = gi;
            END_LANES
        }
        // through the normalisation: in_q = (q - mean) / std ; in_v = v / vn ; h = scale * out - vn
        FOR_LANES
        if (lane < NX) {
            double gd = lane < NQ ? net.scale * gin[lane] / net.stdv : net.scale * gin[lane] / vn;
            if (!clip && lane >= net.vstart) {
                // through vn = |x[vstart:]|: the velocity inputs v / vn and the -vn term
                double dot = 0.0;                // sum_m gin_v[m] v_m
#pragma unroll
                for (int m = 0; m < NQ; ++m) dot += gin[NQ + m] * x[NQ + m];
                gd -= (net.scale * dot / (vn * vn) + 1.0) * x[lane] / vn;
            }
            grad[lane] = gd;
        }
        END_LANES
    }
    UNIFORM_SYNC();
    return h;
}

}  // namespace vboc
