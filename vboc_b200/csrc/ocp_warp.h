// ocp_warp.h -- one warp solves one OCP: SQP with an L1-merit line search around a Mehrotra
// predictor-corrector interior-point method whose Newton systems are solved by a Riccati
// recursion over the shooting stages.
//
// Stands in for acados' ocp_nlp_sqp / ocp_nlp_sqp_rti + HPIPM + BLASFEO + the generated model code
// behind `self.ocp_solver.solve()` (VBOC/triplependulum_class_vboc.py:189,
// AL/triplependulum_class_al.py:162).  Written in the lane-region style of warp_spmd.h.
//
// Mapping of the work onto the 32 lanes:
//   * stage-parallel work (ERK4 + tangents, residuals, step lengths, updates, merit function) runs
//     as flat lane-strided loops over (stage, component) pairs -- coalesced on the stage-major
//     workspace;
//   * the Riccati recursion is serial in the stage index; per stage the (nu+nx)^2 block algebra
//     is spread over the lanes through shared memory: P+[B A] (nz*nx dot products), the symmetric
//     M = H + [B A]'P+[B A] (nz(nz+1)/2), and the Schur complement P = Mxx - Lxu Lxu' (nx*nx) with the
//     tiny nu x nu Cholesky done redundantly in every lane's registers (no communication).
//   * the dt state of the VBOC models is pinned by the reference (lb = ub = dt_sym at every stage,
//     VBOC/triplependulum_vboc.py:98-103) and is eliminated: nx = 2n, the RK4 step is h = dt.
//   * equalities are eliminated exactly instead of being handed to the IPM as lb == ub pairs:
//     stage 0  x_0 = c0 + Z0 y  (fixed components, v_0 parallel to d), terminal v_N = const through
//     the last control.  See DESIGN.md.
#pragma once
#include <stddef.h>

#include "../../include/vboc_b200.h"
#include "dynamics.h"
#include "nn_margin.h"
#include "warp_spmd.h"

namespace vboc {

// Kernel-internal family: the VBOC OCP with the Cartesian path constraint of
// VBOC/Cartesian constraints/doublependulum_class_fixedveldir.py:154-160 as a hard row at stages 0..N-1
// (vboc_set_cartesian on a VBOC-family handle; n = 2)
constexpr int VBOC_FAMILY_CART = 3;

// one problem, reference-shaped device arrays
struct Prob {
    int N;
    int nxr;     // row length of xg / x (2n+1 for the VBOC family, 2n for AL)
    double h;    // RK4 step (pinned dt, or Tf/N)
    double wt;   // VBOC: weight of dt in the cost (p[n]); contributes wt*h*N to get_cost()
    const double *xg, *ug, *p;
    const double *lbx0, *ubx0, *lbx, *ubx, *lbxN, *ubxN, *lbu, *ubu;
    const double *dir;  // unit direction d of the stage-0 constraint (I - d d')v_0 = 0, or nullptr
    double *x, *u;
    vboc_stats *st;
    // optional export of the KKT multipliers at the returned iterate (vboc_download_multipliers):
    // pi [N][2n] of the shooting equalities, lam [N+1][3n][2] of the (lower, upper) bounds on z = [u; q; v]
    double *pi_out = nullptr, *lam_out = nullptr;
    // MPC family (VBOC/Safe MPC/hard_terminal_constraints/doublependulum_class_fixedveldir.py:123-212): diagonal
    // LINEAR_LS weights in z = [u; x] ordering (stage: W = blkdiag(Q, R), terminal: W_e = Q), the reference y_ref of this
    // problem, and the terminal constraint lh <= h(x_N) <= uh with h the learned margin (nn_margin.h)
    const double *Wz = nullptr, *WzN = nullptr, *yref = nullptr, *yrefN = nullptr;
    const NnNet *nn = nullptr;
    double lh = 0.0, uh = 0.0;
    double *lamg_out = nullptr;  // [2] multipliers of the terminal constraint at the returned iterate (optional)
    // soft rows (VBOC/Safe MPC/{parallel,receiding_hard_constraints,soft_traj_constraints}/doublependulum_class_fixedveldir.py:
    // con_h_expr = con_h_expr_e, idxsh = idxsh_e = [0]): the margin row at EVERY stage 0..N, both sides softened by slacks
    // sl, su >= 0 with the per-stage penalties of cost_set(i, "Zl", ...) -- rowZ [N+1][4] = (Zl, Zu, zl, zu), or nullptr
    // for all zero.  rows_soft = 0: the hard terminal row only.
    int rows_soft = 0;
    const double *rowZ = nullptr;
    double cart_xc = 0.0, cart_yc = 0.0;  // CART: centre of the circle the end effector must stay outside of
    double *rowm_out = nullptr;  // [N+1][6] (lam_l, lam_u, lam_sl, lam_su, sl, su) of the rows at the returned iterate
    // AL family: the guess network evaluated inside the kernel (compute_problem_nnguess); the state guess is then
    // computed from the initial state instead of read from xg, and optionally exported (reference-shaped rows)
    const GuessNet *gnn = nullptr;
    double *xg_out = nullptr;
};

template <int NQ>
struct Dim {
    static constexpr int NX = 2 * NQ, NU = NQ, NZ = 3 * NQ, NC = 2 * NZ;
    static constexpr int FS = NU * NU + NX * NU + NX * NX;  // Riccati factor record per stage
    static constexpr int MFS = NZ * NZ + NZ + NU;           // last-stage record
};

// Row record of the MPC family: one general two-sided constraint  lh <= h(x_k) + sl_k,  h(x_k) - su_k <= uh  per stage
// (the learned margin), optionally softened; field offsets in doubles.  IPM order of its one-sided constraints:
// 0 row lower, 1 row upper, 2 slack sl >= 0, 3 slack su >= 0.
struct RowF {
    static constexpr int GC = 0;     // [6] gradient of h at the NLP iterate
    static constexpr int GH = 6, LGD = 7, UGD = 8;            // h, bounds of the linearised row (lh - h, uh - h)
    static constexpr int WEFF = 9, BARG = 10, Q1G = 11, Q2G = 12;  // barrier Hessian weight / gradient pieces (slacks eliminated)
    static constexpr int ZS = 13;    // [4] Zl, Zu, zl, zu
    static constexpr int ON = 17;    // 1.0: the row exists at this stage
    static constexpr int LAMG = 18, LAMS = 20, SL = 22;       // NLP: row multipliers, slack-bound multipliers, slacks
    static constexpr int WGM = 24, WSM = 26;                  // merit weights
    static constexpr int LQ = 28, TQ = 32;                    // [4] QP multipliers / slacks of the inequalities
    static constexpr int SIG = 36, DSIG = 38;                 // [2] QP value of (sl, su) and its step
    static constexpr int RD = 40, PR = 44;                    // [4] bound residuals, second-order products
    static constexpr int SIZE = 48;
};

// Per-stage record: everything the serial Riccati sweeps read at stage k, contiguous so that one
// TMA bulk copy brings it into the shared-memory ring.
template <int NQ>
struct Rec {
    using D = Dim<NQ>;
#if VB_REC_WINDOWS
    // Field order: each sweep reads one contiguous window of the record, so its TMA copy moves only that window.
    static constexpr int HH = 0;                       // effective Hessian diagonal  hd + lam/t (both sides)
    static constexpr int RR = HH + D::NZ;              // predictor gradient
    static constexpr int RB = RR + D::NZ;              // dynamics residual of the QP iterate (even offset)
    static constexpr int BAT = RB + D::NX;             // [B A]' (nz x nx): column j of [B A] contiguous (even offset)
    static constexpr int LUU = BAT + D::NZ * D::NX;    // nu x nu Cholesky factor, inverse diagonal
    static constexpr int LXU = LUU + D::NU * D::NU;    // nx x nu
    static constexpr int YV = LXU + D::NX * D::NU;     // Luu^-1 m_u
    static constexpr int MB = YV + D::NU;              // [B A]' P+ beta (kept for the re-solves)
    static constexpr int Q1 = MB + D::NZ;              // corrector gradient = Q1 - sigma*mu*Q2
    static constexpr int Q2 = Q1 + D::NZ;
    static constexpr int END = Q2 + D::NZ;
    // windows (doubles, even bounds = 16-byte aligned): factorisation, forward sweep, re-solve
    static constexpr int W_FAC0 = 0, W_FAC1 = LUU;
    static constexpr int W_FWD0 = RB, W_FWD1 = MB;
    static constexpr int W_SOL0 = RR & ~1, W_SOL1 = (END + 1) & ~1;
    static_assert(RB % 2 == 0 && BAT % 2 == 0 && LUU % 2 == 0 && MB % 2 == 0, "16-byte aligned windows");
#else
    static constexpr int BAT = 0;                      // [B A]' (nz x nx): column j of [B A] contiguous
    static constexpr int RB = BAT + D::NZ * D::NX;     // dynamics residual of the QP iterate
    static constexpr int HH = RB + D::NX;              // effective Hessian diagonal  hd + lam/t (both sides)
    static constexpr int RR = HH + D::NZ;              // predictor gradient
    static constexpr int Q1 = RR + D::NZ;              // corrector gradient = Q1 - sigma*mu*Q2
    static constexpr int Q2 = Q1 + D::NZ;
    static constexpr int MB = Q2 + D::NZ;              // [B A]' P+ beta (kept for the re-solves)
    static constexpr int LUU = MB + D::NZ;             // nu x nu Cholesky factor, inverse diagonal
    static constexpr int LXU = LUU + D::NU * D::NU;    // nx x nu
    static constexpr int YV = LXU + D::NX * D::NU;     // Luu^-1 m_u
    static constexpr int END = YV + D::NU;
    static constexpr int W_FAC0 = 0, W_FAC1 = (END + 1) & ~1, W_FWD0 = 0, W_FWD1 = W_FAC1, W_SOL0 = 0, W_SOL1 = W_FAC1;
#endif
    static constexpr int SIZE = (END + 1) & ~1;        // doubles; bytes are a multiple of 16
    static constexpr int BYTES = SIZE * 8;
};

#ifndef VB_RING_DEPTH
#define VB_RING_DEPTH 3  // measured: 3 slots leave the SM the 164 KB shared-memory carve-out (92 KB of L1) and beat 4 and 2
#endif
constexpr int RING_DEPTH = VB_RING_DEPTH;

// dot product of two contiguous, 16-byte aligned vectors of even length: 128-bit shared-memory loads
// (LDS.128) and two accumulators
template <int LEN>
VB_HD double dotv(const double *a, const double *b, double init = 0.0) {
    static_assert(LEN % 2 == 0, "even length");
#if defined(__CUDA_ARCH__)
    const double2 *a2 = reinterpret_cast<const double2 *>(a), *b2 = reinterpret_cast<const double2 *>(b);
    double a0 = init, a1 = 0.0;
#pragma unroll
    for (int m = 0; m < LEN / 2; ++m) {
        const double2 x = a2[m], y = b2[m];
        a0 += x.x * y.x, a1 += x.y * y.y;
    }
    return a0 + a1;
#else
    double a0 = init, a1 = 0.0;
    for (int m = 0; m < LEN; m += 2) a0 += a[m] * b[m], a1 += a[m + 1] * b[m + 1];
    return a0 + a1;
#endif
}
#ifndef VB_CON_ROLLED
#define VB_CON_ROLLED 1  // the two sides of a bound as a rolled loop: half the body, +3 % (instruction fetch bound)
#endif
#ifndef VB_PF_RES
#define VB_PF_RES 0  // measured: the residual pass is no faster with it, the constraint and update passes are
#endif
#ifndef VB_PF_CON
#define VB_PF_CON 1
#endif
#ifndef VB_PF_UPD
#define VB_PF_UPD 1
#endif
#ifndef VB_STORE_RMB
#define VB_STORE_RMB 0  // lam*t of the constraint pass: recomputed (two loads and two stores fewer per element, +1.7 %)
#endif
#ifndef VB_REC_WINDOWS
#define VB_REC_WINDOWS 1  // each sweep's TMA copy moves only the fields it reads (-12 % DRAM bytes, +2 %)
#endif
#ifndef VB_RG_ALL
#define VB_RG_ALL 0  // the stationarity residual is stored for stage 0 only (the only one read back)
#endif
#ifndef VB_VEC_RES
#define VB_VEC_RES 1  // 128-bit loads of the [B A] column and pi in the residual pass (+0.9 %)
#endif
#ifndef VB_PAIR_LAYOUT
#define VB_PAIR_LAYOUT 1  // lower / upper entries of a bound adjacent: one 128-bit load per array in the flat passes (+1.8 %)
#endif
#ifndef VB_RECOMPUTE_STEP
#define VB_RECOMPUTE_STEP 1  // slack / multiplier steps recomputed inside the fused update: DT / DLAM never written (+0.9 %)
#endif
#ifndef VB_FLAT_UNROLL
#define VB_FLAT_UNROLL 1  // unroll factor of the three flat passes of the IPM (tuning: 2 needs ~128 registers)
#endif
constexpr int FLAT_UNROLL = VB_FLAT_UNROLL;
#ifndef VB_PF_DIST
#define VB_PF_DIST 64  // software prefetch distance of the flat passes: two lane-strided iterations ahead
#endif

// dot product with strides and two accumulators (halves the dependent FP64 chain of the sweeps)
template <int LEN>
VB_HD double dot2(const double *a, int sa, const double *b, int sb, double init = 0.0) {
    double a0 = init, a1 = 0.0;
#pragma unroll
    for (int m = 0; m + 1 < LEN; m += 2) {
        a0 += a[m * sa] * b[m * sb];
        a1 += a[(m + 1) * sa] * b[(m + 1) * sb];
    }
    if (LEN & 1) a0 += a[(LEN - 1) * sa] * b[(LEN - 1) * sb];
    return a0 + a1;
}

// Global-memory workspace of one warp slot, stage-major.
template <int NQ>
struct Work {
    using D = Dim<NQ>;
    using R = Rec<NQ>;
    static constexpr int PPS = D::NX * D::NX + D::NX;  // value function record: P (nx x nx), p (nx)
    double *SR;                      // stage records (16-byte aligned)
    double *Z, *PI, *LAM, *BD;       // NLP iterate z_k = [u_k; x_k], multipliers, shooting gaps
    double *DZ, *PIQ, *LAMQ, *TQ;    // QP iterate
    double *DV, *DLAM, *DT;          // IPM step
    double *RG, *RD, *RM, *RMB;      // IPM residuals (RMB = lam*t, RM = dt_aff*dlam_aff)
    double *PP, *MF;                 // value functions; last-stage record
    double *WDYN, *WB, *ZT;          // merit weights, trial point
    double *NNA;                     // MPC family: activations of the margin network (2 x NN_HMAX)
    double *ROW;                     // MPC family: row records (SMAX x RowF::SIZE), behind TOTAL (doubles_rows())
    // The workspace is carved with the compile-time stride SMAX so that every array is the slot base
    // plus a constant (no pointer table in registers, immediate offsets in the load/store
    // instructions); N_max <= SMAX - 1 is checked by vboc_create.
    static constexpr size_t SMAX = 130;  // even: every array then starts on a 16-byte boundary (128-bit loads)
    static constexpr size_t O_SR = 0;
    static constexpr size_t O_Z = O_SR + SMAX * R::SIZE, O_PI = O_Z + SMAX * D::NZ, O_LAM = O_PI + SMAX * D::NX;
    static constexpr size_t O_BD = O_LAM + SMAX * D::NC, O_DZ = O_BD + SMAX * D::NX, O_PIQ = O_DZ + SMAX * D::NZ;
    static constexpr size_t O_LAMQ = O_PIQ + SMAX * D::NX, O_TQ = O_LAMQ + SMAX * D::NC, O_DV = O_TQ + SMAX * D::NC;
    static constexpr size_t O_DLAM = O_DV + SMAX * D::NZ, O_DT = O_DLAM + SMAX * D::NC, O_RG = O_DT + SMAX * D::NC;
    static constexpr size_t O_RD = O_RG + SMAX * D::NZ, O_RM = O_RD + SMAX * D::NC, O_RMB = O_RM + SMAX * D::NC;
    static constexpr size_t O_PP = O_RMB + SMAX * D::NC, O_WDYN = O_PP + SMAX * PPS, O_WB = O_WDYN + SMAX * D::NX;
    static constexpr size_t O_ZT = O_WB + SMAX * D::NC, O_MF = O_ZT + SMAX * D::NZ;
    static constexpr size_t O_NNA = (O_MF + D::MFS + 1) & ~(size_t)1;
    static constexpr size_t TOTAL = O_NNA + 2 * NN_HMAX;
    static_assert(O_PIQ % 2 == 0 && O_LAMQ % 2 == 0 && O_TQ % 2 == 0 && O_DLAM % 2 == 0 && O_DT % 2 == 0 && O_RD % 2 == 0 &&
                      O_RM % 2 == 0, "16-byte aligned arrays");
    static VB_HD size_t doubles(int) { return TOTAL; }
    static VB_HD size_t doubles_rows(int) { return TOTAL + SMAX * RowF::SIZE; }  // MPC family
    VB_HD void carve(double *b, int) {
        SR = b + O_SR;
        Z = b + O_Z, PI = b + O_PI, LAM = b + O_LAM, BD = b + O_BD;
        DZ = b + O_DZ, PIQ = b + O_PIQ, LAMQ = b + O_LAMQ, TQ = b + O_TQ;
        DV = b + O_DV, DLAM = b + O_DLAM, DT = b + O_DT;
        RG = b + O_RG, RD = b + O_RD, RM = b + O_RM, RMB = b + O_RMB;
        PP = b + O_PP, WDYN = b + O_WDYN, WB = b + O_WB, ZT = b + O_ZT, MF = b + O_MF, NNA = b + O_NNA;
        ROW = b + TOTAL;
    }
};

// Shared-memory block of one warp.
template <int NQ>
struct alignas(16) Smem {
    using D = Dim<NQ>;
    double ring[RING_DEPTH][Rec<NQ>::SIZE];  // TMA destination: stage records
    unsigned long long bar[RING_DEPTH];      // one mbarrier per ring slot
    // problem constants
    double lb[3][D::NZ], ub[3][D::NZ];  // stage classes 0, 1..N-1, N in z ordering
    double Z0[D::NX][D::NX];            // orthonormal basis of the stage-0 free subspace, zero padded
    double c0[D::NX], cN[D::NX], w[NQ];
    double h, wtdt;
    int N, fixed0, fixedN, termfix, nact;
    unsigned char tri[64];  // (row << 4 | column) of the entries of the lower triangle of M, row-major
    // Riccati staging
    alignas(16) double P[D::NX][D::NX];
    alignas(16) double PBAT[D::NZ][D::NX];
    alignas(16) double tb[D::NX];
    alignas(16) double pvec[D::NX];
    double M[D::NZ][D::NZ], m[D::NZ], dx[2][D::NX];
    double e0[D::NX], eN[D::NX], hhN[D::NX], rN[D::NX], nuv[D::NX], Lz[D::NX][D::NX], dzi[D::NX], Pe[D::NX];
    double K[D::NU][D::NX], tz[D::NZ], va[D::NX], vb[D::NX], vc[D::NX], vt[D::NX];
    // merit weights / multipliers of the eliminated equalities
    double w0[D::NX], wN[D::NX], nu0q[D::NX], nuNq[D::NX];
};

// Extra shared-memory block of one warp for the MPC family only (the other families' kernels do not carry it):
// tracking cost and the terminal constraint row  lgd <= gc' dx_N <= ugd  of the QP.
template <int NQ>
struct alignas(16) SmemMpc {
    using D = Dim<NQ>;
    double Wz[D::NZ], WzN[D::NX], yref[D::NZ], yrefN[D::NX];
    double gc[D::NX], gx[D::NX], nnio[2 * D::NX + 2];  // scratch of the margin evaluation
    double lh, uh;
    double xc, yc;  // CART: circle centre
    int soft;   // rows at every stage with slacks (Prob::rows_soft), else the hard terminal row
};

template <int NQ, int FAM>
struct WarpSolver {
    using D = Dim<NQ>;
    static constexpr int NX = D::NX, NU = D::NU, NZ = D::NZ, NC = D::NC, FS = D::FS;
    static constexpr int OFF_LXU = NU * NU, OFF_P = NU * NU + NX * NU;
    static constexpr int TRI = NZ * (NZ + 1) / 2;

    using R = Rec<NQ>;
    Smem<NQ> &s;
    Work<NQ> w;
    const vboc_opts &o;
    VbRing ring;
    SmemMpc<NQ> *g = nullptr;    // MPC family only
    const NnNet *nn = nullptr;

    VB_DEV WarpSolver(Smem<NQ> &s_, const Work<NQ> &w_, const vboc_opts &o_, SmemMpc<NQ> *g_ = nullptr)
        : s(s_), w(w_), o(o_), g(g_) {
        RING_INIT(ring, s.bar, RING_DEPTH);
    }
    static constexpr bool MPC = FAM == VBOC_FAMILY_MPC;
    static constexpr bool CART = FAM == VBOC_FAMILY_CART;
    static constexpr bool ROWS = MPC || CART;               // families that carry general rows (RowF records)
    static constexpr bool VBOCLIKE = FAM == VBOC_FAMILY_VBOC || CART;  // linear cost on v_0, dt eliminated
    // row record of stage k (MPC family)
    VB_DEV double *row(int k) const { return w.ROW + (size_t)k * RowF::SIZE; }
    VB_DEV int row_first() const { return (CART || g->soft) ? 0 : s.N; }  // first / last stage that carries a row
    VB_DEV int row_last() const { return CART ? s.N - 1 : s.N; }

    // index of the (k, i, side) bound constraint in the constraint arrays (LAM, LAMQ, TQ, DLAM, DT, RD, RM, RMB, WB).
    // VB_PAIR_LAYOUT: the lower / upper entries of a component are adjacent, so both come with one 128-bit load.
    static VB_DEV int CI(int k, int i, int sd) {
#if VB_PAIR_LAYOUT
        return (k * NZ + i) * 2 + sd;
#else
        return k * NC + sd * NZ + i;
#endif
    }
    // ---------------------------------------------------------------- problem structure helpers
    VB_DEV int sclass(int k) const { return k == 0 ? 0 : (k == s.N ? 2 : 1); }
    VB_DEV bool active(int k, int i) const {
        if (k == s.N) {
            if (i < NU) return false;
            if ((s.fixedN >> (i - NU)) & 1) return false;
        }
        if (k == 0 && i >= NU && ((s.fixed0 >> (i - NU)) & 1)) return false;
        return true;
    }
    // gradient / Hessian diagonal of the stage cost (VBOC: EXTERNAL linear cost
    // VBOC/triplependulum_class_vboc.py:82-87 + levenberg_marquardt; AL: LINEAR_LS on the
    // velocities, AL/triplependulum_class_al.py:98-115, Gauss-Newton)
    VB_DEV double cost_g(int k, int i, double zval) const {
        if (VBOCLIKE) return (k == 0 && i >= NU + NQ) ? s.w[i - NU - NQ] : 0.0;
        if constexpr (MPC) {
            // LINEAR_LS, Gauss-Newton: the stage costs are scaled by the time step, the terminal one is not
            if (k < s.N) return s.h * g->Wz[i] * (zval - g->yref[i]);
            return i >= NU ? g->WzN[i - NU] * (zval - g->yrefN[i - NU]) : 0.0;
        }
        return (i >= NU + NQ) ? 2.0 * (k < s.N ? s.h : 1.0) * zval : 0.0;
    }
    VB_DEV double cost_h(int k, int i) const {
        double hd = o.levenberg_marquardt;
        if (FAM == VBOC_FAMILY_AL && i >= NU + NQ) hd += 2.0 * (k < s.N ? s.h : 1.0);
        if constexpr (MPC) hd += k < s.N ? s.h * g->Wz[i] : (i >= NU ? g->WzN[i - NU] : 0.0);
        return hd;
    }
    // e = (I - Z0 Z0')(x - c0): violation of the stage-0 equalities; x, e in shared memory
    VB_DEV void eq0_violation(const double *x, double *e) {
        FOR_LANES
        if (lane < NX) {
            double a = 0.0;
#pragma unroll
            for (int i = 0; i < NX; ++i) a += s.Z0[i][lane] * (x[i] - s.c0[i]);
            s.vt[lane] = a;
        }
        END_LANES
        FOR_LANES
        if (lane < NX) {
            double a = x[lane] - s.c0[lane];
#pragma unroll
            for (int c = 0; c < NX; ++c) a -= s.Z0[lane][c] * s.vt[c];
            e[lane] = a;
        }
        END_LANES
    }
    // v <- Z0 Z0' v; v in shared memory
    VB_DEV void proj0(double *v) {
        FOR_LANES
        if (lane < NX) {
            double a = 0.0;
#pragma unroll
            for (int i = 0; i < NX; ++i) a += s.Z0[i][lane] * v[i];
            s.vt[lane] = a;
        }
        END_LANES
        FOR_LANES
        if (lane < NX) {
            double a = 0.0;
#pragma unroll
            for (int c = 0; c < NX; ++c) a += s.Z0[lane][c] * s.vt[c];
            v[lane] = a;
        }
        END_LANES
    }

    // ---------------------------------------------------------------- problem load / store
    VB_DEV void load_problem(const Prob &pb) {
        const int N = pb.N;
        FOR_LANES
        if (lane == 0) {
            s.N = N, s.h = pb.h;
            s.wtdt = VBOCLIKE ? pb.wt * pb.h * N : 0.0;
            int f0 = 0, fN = 0;
            for (int i = 0; i < NX; ++i) {
                bool a = pb.lbx0[i] == pb.ubx0[i], b = pb.lbxN[i] == pb.ubxN[i];
                f0 |= (int)a << i, fN |= (int)b << i;
                s.c0[i] = a ? pb.lbx0[i] : 0.0;
                s.cN[i] = b ? pb.lbxN[i] : 0.0;
                for (int c = 0; c < NX; ++c) s.Z0[i][c] = 0.0;
                s.w0[i] = s.wN[i] = 0.0;
            }
            int ny = 0;
            for (int i = 0; i < NX; ++i) {
                if ((f0 >> i) & 1) continue;
                if (pb.dir && i >= NQ) continue;
                s.Z0[i][ny++] = 1.0;
            }
            if (pb.dir) {
                for (int i = 0; i < NQ; ++i) s.Z0[NQ + i][ny] = pb.dir[i];
                ++ny;
            }
            int nf0 = 0, nfN = 0;
            for (int i = 0; i < NX; ++i) nf0 += (f0 >> i) & 1, nfN += (fN >> i) & 1;
            s.fixed0 = f0, s.fixedN = fN, s.termfix = nfN != 0;
            s.nact = (N + 1) * NZ - NU - nf0 - nfN;
            for (int i = 0; i < NQ; ++i) s.w[i] = VBOCLIKE ? pb.p[i] : 0.0;
            if constexpr (MPC) {
                // hard terminal row: one more two-sided constraint; soft rows: per stage the row (two sides) and the two
                // slack bounds, i.e. two two-sided constraints' worth
                s.nact += pb.rows_soft ? 2 * (N + 1) : 1;
                for (int i = 0; i < NZ; ++i) g->Wz[i] = pb.Wz[i], g->yref[i] = pb.yref[i];
                for (int i = 0; i < NX; ++i) g->WzN[i] = pb.WzN[i], g->yrefN[i] = pb.yrefN[i];
                g->lh = pb.lh, g->uh = pb.uh, g->soft = pb.rows_soft;
            }
            if constexpr (CART) {
                s.nact += N;  // one hard two-sided row per stage 0..N-1
                g->lh = pb.lh, g->uh = pb.uh, g->soft = 0, g->xc = pb.cart_xc, g->yc = pb.cart_yc;
            }
        }
        if constexpr (ROWS) {
            if constexpr (MPC) nn = pb.nn;
            // acados reset(): zero row multipliers and slacks
            for (int k = lane; k <= N; k += 32) {
                double *rw = w.ROW + (size_t)k * RowF::SIZE;
#pragma unroll 1
                for (int j = 0; j < RowF::SIZE; ++j) rw[j] = 0.0;
                rw[RowF::ON] = (CART ? k < N : (pb.rows_soft || k == N)) ? 1.0 : 0.0;
                if (pb.rows_soft && pb.rowZ) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) rw[RowF::ZS + j] = pb.rowZ[k * 4 + j];
                }
            }
        }
        static_assert(TRI <= 64 && NZ <= 16, "triangle index table");
        for (int idx = lane; idx < TRI; idx += 32) {
            int a_ = 0;
            while ((a_ + 1) * (a_ + 2) / 2 <= idx) ++a_;
            s.tri[idx] = (unsigned char)((a_ << 4) | (idx - a_ * (a_ + 1) / 2));
        }
        for (int idx = lane; idx < 3 * NZ; idx += 32) {
            int sc = idx / NZ, i = idx - sc * NZ;
            const double *l = sc == 0 ? pb.lbx0 : (sc == 1 ? pb.lbx : pb.lbxN);
            const double *u = sc == 0 ? pb.ubx0 : (sc == 1 ? pb.ubx : pb.ubxN);
            s.lb[sc][i] = i < NU ? pb.lbu[i] : l[i - NU];
            s.ub[sc][i] = i < NU ? pb.ubu[i] : u[i - NU];
        }
        for (int idx = lane; idx < (N + 1) * NZ; idx += 32) {
            int k = idx / NZ, i = idx - k * NZ;
            w.Z[idx] = i < NU ? (k < N ? pb.ug[k * NU + i] : 0.0) : pb.xg[(size_t)k * pb.nxr + i - NU];
        }
        // acados reset(): zero multipliers
#pragma unroll 1
        for (int idx = lane; idx < N * NX; idx += 32) w.PI[idx] = 0.0;
#pragma unroll 1
        for (int idx = lane; idx < (N + 1) * NC; idx += 32) w.LAM[idx] = 0.0;
        END_LANES
        if constexpr (FAM == VBOC_FAMILY_AL) {
            if (pb.gnn) {
                // compute_problem_nnguess (AL/triplependulum_class_al.py:171-201): stage 0 is the initial state itself,
                // stages 1..N the de-normalised output of the guess network; controls stay at their (zero) guess
                FOR_LANES
                if (lane < NX) s.va[lane] = pb.lbx0[lane], w.Z[NU + lane] = pb.lbx0[lane];
                END_LANES
                double *Z = w.Z;
                guess_forward<NQ>(*pb.gnn, s.va, w.NNA, w.NNA + NN_HMAX, [Z](int j, double v) {
                    const int k = j / NX + 1, i = j - (k - 1) * NX;
                    Z[k * NZ + NU + i] = v;
                });
                if (pb.xg_out) {
                    FOR_LANES
                    for (int idx = lane; idx < (N + 1) * NX; idx += 32) {
                        const int k = idx / NX, i = idx - k * NX;
                        pb.xg_out[(size_t)k * pb.nxr + i] = w.Z[k * NZ + NU + i];
                    }
                    END_LANES
                }
            }
        }
    }

    VB_DEV void store_solution(const Prob &pb, const vboc_stats &st) {
        const int N = s.N;
        FOR_LANES
        for (int idx = lane; idx < (N + 1) * pb.nxr; idx += 32) {
            int k = idx / pb.nxr, i = idx - k * pb.nxr;
            pb.x[idx] = i < NX ? w.Z[k * NZ + NU + i] : s.h;
        }
        for (int idx = lane; idx < N * NU; idx += 32) {
            int k = idx / NU, i = idx - k * NU;
            pb.u[idx] = w.Z[k * NZ + i];
        }
        if (lane == 0) *pb.st = st;
        if constexpr (ROWS) {
            if (pb.lamg_out && lane < 2) pb.lamg_out[lane] = row(N)[RowF::LAMG + lane];
            if (pb.rowm_out) {
                for (int idx = lane; idx < (N + 1) * 6; idx += 32) {
                    const int k = idx / 6, j = idx - k * 6;
                    pb.rowm_out[idx] = row(k)[RowF::LAMG + j];  // LAMG[2], LAMS[2], SL[2] are adjacent
                }
            }
        }
        if (pb.pi_out) {
            for (int idx = lane; idx < N * NX; idx += 32) pb.pi_out[idx] = w.PI[idx];
            for (int idx = lane; idx < (N + 1) * NZ; idx += 32) {
                int k = idx / NZ, i = idx - k * NZ;
                pb.lam_out[2 * idx] = w.LAM[CI(k, i, 0)], pb.lam_out[2 * idx + 1] = w.LAM[CI(k, i, 1)];
            }
        }
        END_LANES
    }

    // ---------------------------------------------------------------- stage-record access
    VB_DEV double *rec(int k) const { return w.SR + (size_t)k * R::SIZE; }

    // ---------------------------------------------------------------- linearisation
    // ERK4 with forward tangents: NZ+1 lanes per shooting interval, lane `dir` < NZ carries the
    // tangent d/dz_dir (one column of [B A]), lane NZ the value (gap).  32/(NZ+1) intervals per pass.
    VB_DEV void linearize() {
        const int N = s.N;
        constexpr int LPS = NZ + 1, SPP = 32 / LPS;
        FOR_LANES
        const int sub = lane / LPS, dir = lane - sub * LPS;
        if (sub < SPP) {
            for (int k = sub; k < N; k += SPP) {
                const double *z = w.Z + (size_t)k * NZ;
                Dual1 x[NX], u[NU], xn[NX];
#pragma unroll
                for (int i = 0; i < NU; ++i) u[i] = {z[i], dir == i ? 1.0 : 0.0};
#pragma unroll
                for (int i = 0; i < NX; ++i) x[i] = {z[NU + i], dir == NU + i ? 1.0 : 0.0};
                rk4_step<NQ, Dual1>(x, u, s.h, xn);
                if (dir < NZ) {
                    double *col = rec(k) + R::BAT + dir * NX;
#pragma unroll
                    for (int i = 0; i < NX; ++i) col[i] = xn[i].d;
                } else {
                    const double *zn = w.Z + (size_t)(k + 1) * NZ + NU;
#pragma unroll
                    for (int i = 0; i < NX; ++i) w.BD[k * NX + i] = xn[i].v - zn[i];
                }
            }
        }
        END_LANES
    }

    // MPC family: h(x_k) and its gradient at the NLP iterate for every stage that carries a row -> the row
    // lh - h <= gc' dx_k (+ sl) ,  gc' dx_k (- su) <= uh - h  of the QP
    // the row function h(x) and (grad != nullptr) its gradient: MPC the learned margin, CART the squared distance of
    // the end effector from the circle centre, (l1 sin q1 + l2 sin q2 - xc)^2 + (l1 cos q1 + l2 cos q2 - yc)^2
    // (VBOC/Cartesian constraints/doublependulum_class_fixedveldir.py:154-156).  x, grad in shared memory; uniform result.
    VB_DEV double row_eval(const double *x, double *grad) {
        if constexpr (MPC) {
            return nn_margin<NQ>(*nn, x, grad, w.NNA, w.NNA + NN_HMAX, g->nnio);
        } else {
            static_assert(!CART || NQ == 2, "the Cartesian constraint is the double pendulum's");
            const double l1 = PendN::l, l2 = PendN::l;  // both links 0.8 m
            double s1, c1, s2, c2;
            sincos(x[0], &s1, &c1);
            sincos(x[1], &s2, &c2);
            const double ex = l1 * s1 + l2 * s2 - g->xc, ey = l1 * c1 + l2 * c2 - g->yc;
            UNIFORM_SYNC();
            if (grad) {
                FOR_LANES
                if (lane < NX) grad[lane] = lane == 0 ? 2.0 * l1 * (ex * c1 - ey * s1) : (lane == 1 ? 2.0 * l2 * (ex * c2 - ey * s2) : 0.0);
                END_LANES
            }
            return ex * ex + ey * ey;
        }
    }
    VB_DEV void linearize_rows() {
        if constexpr (ROWS) {
            const int N = s.N;
            (void)N;
#pragma unroll 1
            for (int k = row_first(); k <= row_last(); ++k) {
                FOR_LANES
                if (lane < NX) g->gx[lane] = w.Z[k * NZ + NU + lane];
                END_LANES
                const double h = row_eval(g->gx, g->gc);
                FOR_LANES
                double *rw = row(k);
                if (lane < NX) rw[RowF::GC + lane] = g->gc[lane];
                if (lane == 0) rw[RowF::GH] = h, rw[RowF::LGD] = g->lh - h, rw[RowF::UGD] = g->uh - h;
                END_LANES
            }
        }
    }
    // h at stage k of a trial point (merit function)
    VB_DEV double row_value(const double *Zs, int k) {
        FOR_LANES
        if (lane < NX) g->gx[lane] = Zs[k * NZ + NU + lane];
        END_LANES
        return row_eval(g->gx, nullptr);
    }
    // cost of the slacks  sum_k 1/2 Z s^2 + z s ; trial: at the trial slacks of the line search (RowF::RD), else at SL
    VB_DEV double slack_cost(bool trial) {
        double c = 0.0;
        if constexpr (ROWS) {
            if (g->soft) {
                const int N = s.N;
                LV(double, acc);
                FOR_LANES
                double a = 0.0;
                for (int k = lane; k <= N; k += 32) {
                    const double *rw = row(k);
#pragma unroll
                    for (int sd = 0; sd < 2; ++sd) {
                        const double sv = trial ? rw[RowF::RD + sd] : rw[RowF::SL + sd];
                        a += (0.5 * rw[RowF::ZS + sd] * sv + rw[RowF::ZS + 2 + sd]) * sv;
                    }
                }
                L(acc) = a;
                END_LANES
                c = WARP_SUM(acc);
            }
        }
        return c;
    }

    // ---------------------------------------------------------------- NLP residuals
    // inf-norms of the Lagrangian gradient, shooting gaps, constraint violation, complementarity
    VB_DEV bool nlp_residuals(double &rs, double &re, double &ri, double &rc) {
        const int N = s.N;
        LV(double, a_s);
        LV(double, a_e);
        LV(double, a_i);
        LV(double, a_c);
        LV(int, bad);
        FOR_LANES
        double vs = 0, ve = 0, vi = 0, vc = 0;
        int nb = 0;
        for (int idx = lane; idx < (N + 1) * NZ; idx += 32) {
            int k = idx / NZ, i = idx - k * NZ;
            double r = 0.0;
            if (!(k == N && i < NU)) {
                double z = w.Z[idx];
                r = cost_g(k, i, z);
                if (k < N) {
                    const double *col = rec(k) + R::BAT + i * NX, *pi = w.PI + k * NX;
#pragma unroll
                    for (int m = 0; m < NX; ++m) r += col[m] * pi[m];
                }
                if (k > 0 && i >= NU) r -= w.PI[(k - 1) * NX + i - NU];
                if constexpr (ROWS) {
                    if (i >= NU) {
                        const double *rw = row(k);
                        if (rw[RowF::ON] != 0.0) r += rw[RowF::GC + i - NU] * (rw[RowF::LAMG + 1] - rw[RowF::LAMG]);
                    }
                }
                if (active(k, i)) {
                    int sc = sclass(k);
                    double ll = w.LAM[CI(k, i, 0)], lu = w.LAM[CI(k, i, 1)];
                    double fl = s.lb[sc][i] - z, fu = z - s.ub[sc][i];
                    r += lu - ll;
                    vi = fmax(vi, fmax(fl, fu));
                    vc = fmax(vc, fmax(fabs(ll * fl), fabs(lu * fu)));
                } else if (k == N) {
                    r = 0.0;
                    vi = fmax(vi, fabs(z - s.cN[i - NU]));
                }
            }
            if (VB_RG_ALL || k == 0) w.RG[idx] = r;  // scratch: the stage-0 state part is projected below
            if (!(k == 0 && i >= NU)) {
                nb |= (r != r);
                vs = fmax(vs, fabs(r));
            }
        }
#pragma unroll 1
        for (int idx = lane; idx < N * NX; idx += 32) {
            double v = w.BD[idx];
            nb |= (v != v);
            ve = fmax(ve, fabs(v));
        }
        if constexpr (ROWS) {
            // the rows  lh <= h(x_k) + sl ,  h(x_k) - su <= uh ,  sl, su >= 0  and the stationarity of the slacks
            for (int k = lane; k <= N; k += 32) {
                const double *rw = row(k);
                if (rw[RowF::ON] == 0.0) continue;
                const double sl = rw[RowF::SL], su = rw[RowF::SL + 1];  // zero for hard rows
                const double fl = g->lh - rw[RowF::GH] - sl, fu = rw[RowF::GH] - su - g->uh;
                vi = fmax(vi, fmax(fl, fu));
                vc = fmax(vc, fmax(fabs(rw[RowF::LAMG] * fl), fabs(rw[RowF::LAMG + 1] * fu)));
                if (g->soft) {
                    vi = fmax(vi, fmax(-sl, -su));
                    vc = fmax(vc, fmax(fabs(rw[RowF::LAMS] * sl), fabs(rw[RowF::LAMS + 1] * su)));
                    const double rsl = rw[RowF::ZS] * sl + rw[RowF::ZS + 2] - rw[RowF::LAMG] - rw[RowF::LAMS];
                    const double rsu = rw[RowF::ZS + 1] * su + rw[RowF::ZS + 3] - rw[RowF::LAMG + 1] - rw[RowF::LAMS + 1];
                    nb |= (rsl != rsl) | (rsu != rsu);
                    vs = fmax(vs, fmax(fabs(rsl), fabs(rsu)));
                }
            }
        }
        L(a_s) = vs, L(a_e) = ve, L(a_i) = vi, L(a_c) = vc, L(bad) = nb;
        END_LANES
        FOR_LANES
        if (lane < NX) s.va[lane] = w.RG[NU + lane], s.vb[lane] = w.Z[NU + lane];
        END_LANES
        proj0(s.va);
        eq0_violation(s.vb, s.vc);
        FOR_LANES
        if (lane < NX) {
            const double va = s.va[lane];
            L(bad) |= (va != va);
            L(a_s) = fmax(L(a_s), fabs(va));
            L(a_i) = fmax(L(a_i), fabs(s.vc[lane]));
        }
        END_LANES
        rs = WARP_MAX(a_s), re = WARP_MAX(a_e), ri = WARP_MAX(a_i), rc = WARP_MAX(a_c);
        bool nan = WARP_ANY(bad);
        UNIFORM_SYNC();
        return !nan;
    }

    // ---------------------------------------------------------------- QP: cold start
    VB_DEV void qp_init() {
        const int N = s.N;
        const double thr0 = 0.1;  // HPIPM cold-start threshold
        FOR_LANES
        for (int idx = lane; idx < (N + 1) * NZ; idx += 32) {
            int k = idx / NZ, i = idx - k * NZ;
            double z = w.Z[idx], v = 0.0, tl = 0, tu = 0, ll = 0, lu = 0;
            if (active(k, i)) {
                int sc = sclass(k);
                double lbd = s.lb[sc][i] - z, ubd = s.ub[sc][i] - z;
                tl = v - lbd, tu = ubd - v;
                if (tl < thr0) {
                    if (tu < thr0) {
                        v = 0.5 * (lbd + ubd);
                        tl = tu = thr0;
                    } else {
                        tl = thr0;
                        v = lbd + thr0;
                    }
                } else if (tu < thr0) {
                    tu = thr0;
                    v = ubd - thr0;
                }
                ll = o.qp_mu0 / tl, lu = o.qp_mu0 / tu;
            } else if (k == N && i >= NU) {
                v = s.cN[i - NU] - z;
            }
            w.DZ[idx] = v;
            w.LAMQ[CI(k, i, 0)] = ll, w.LAMQ[CI(k, i, 1)] = lu;
            w.TQ[CI(k, i, 0)] = tl, w.TQ[CI(k, i, 1)] = tu;
        }
#pragma unroll 1
        for (int idx = lane; idx < N * NX; idx += 32) w.PIQ[idx] = 0.0;
        END_LANES
        FOR_LANES
        if (lane < NX) s.vb[lane] = w.Z[NU + lane] + w.DZ[NU + lane];
        END_LANES
        eq0_violation(s.vb, s.vc);
        FOR_LANES
        if (lane < NX) w.DZ[NU + lane] -= s.vc[lane];
        END_LANES
        if constexpr (ROWS) {
            // HPIPM cold start of a general constraint: slack = distance to the bound at the initial step, not below
            // thr0; a softening slack starts at its bound and is pushed inside by thr0 like a box-bounded variable
            FOR_LANES
            for (int k = lane; k <= N; k += 32) {
                double *rw = row(k);
                if (rw[RowF::ON] == 0.0) continue;
                double g0 = 0.0;
#pragma unroll
                for (int i = 0; i < NX; ++i) g0 += rw[RowF::GC + i] * w.DZ[k * NZ + NU + i];
                rw[RowF::DSIG] = rw[RowF::DSIG + 1] = 0.0;
                if (g->soft) {
                    // A softened row can always start strictly feasible: the slack absorbs the violation (not below thr0,
                    // like a box-bounded variable at its bound).  The row multiplier starts where the slack's own
                    // stationarity Z s + z - lam_row - lam_slack = 0 holds, so that a large penalty does not enter the
                    // first Newton systems as a residual of size Z s (with Z ~ 1e6 ... 1e12 the IPM otherwise spends its
                    // iterations growing that multiplier by a factor per step).
                    const double viol[2] = {rw[RowF::LGD] - g0, g0 - rw[RowF::UGD]};
#pragma unroll
                    for (int sd = 0; sd < 2; ++sd) {
                        const double sg = fmax(viol[sd] + thr0, thr0), t = sg - viol[sd];
                        const double ls = o.qp_mu0 / sg;
                        rw[RowF::SIG + sd] = sg, rw[RowF::TQ + sd] = t, rw[RowF::TQ + 2 + sd] = sg, rw[RowF::LQ + 2 + sd] = ls;
                        rw[RowF::LQ + sd] = fmax(o.qp_mu0 / t, rw[RowF::ZS + sd] * sg + rw[RowF::ZS + 2 + sd] - ls);
                    }
                } else {
                    rw[RowF::SIG] = rw[RowF::SIG + 1] = 0.0;
                    const double tl = fmax(g0 - rw[RowF::LGD], thr0), tu = fmax(rw[RowF::UGD] - g0, thr0);
                    rw[RowF::TQ] = tl, rw[RowF::TQ + 1] = tu;
                    rw[RowF::LQ] = o.qp_mu0 / tl, rw[RowF::LQ + 1] = o.qp_mu0 / tu;
                }
            }
            END_LANES
        }
    }

    // ---------------------------------------------------------------- QP: residuals
    // RG (stationarity), RB (dynamics, into the stage records), RD (bounds), RMB (lam*t), and the
    // Newton-system data of the predictor: HH = hd + lam/t, RR = RG + (lam*t - lam*rd)/t (signed).
    // Returns mu, norms by reference.
    // upd: first apply the step of the previous IPM iteration to DZ, LAMQ, TQ (z += as dz, lam/t = max(. + as d., min))
    // -- the update rides on the loads this pass does anyway instead of two more passes over the arrays.
    VB_DEV double qp_residuals(double &ng, double &nb_, double &nd, double &nm, bool &nan, bool upd = false,
                               double as = 0.0, int umode = 1, double usm = 0.0) {
        const int N = s.N;
        LV(double, a_g);
        LV(double, a_b);
        LV(double, a_d);
        LV(double, a_m);
        LV(double, a_mu);
        LV(int, bad);
        LV(double, a_lm);  // ROWS: largest row multiplier (scale of the stationarity test, see below)
        if constexpr (ROWS) {
            // the rows first (lanes over the stages): their updated multipliers enter the stationarity residual of x_k in
            // the flat loop below, their barrier terms (slacks eliminated) the stage Hessian of the
            // backward sweep.  One-sided constraints of a row: 0 lower, 1 upper, 2 sl >= 0, 3 su >= 0.
            FOR_LANES
            double vg = 0, vd = 0, vm = 0, mu = 0, lmx = 0;
            int nb = 0;
            const int nsd = g->soft ? 4 : 2;
            for (int k = N - lane; k >= 0; k -= 32) {  // lane 0 takes the terminal row (summation order of the hard-row kernel)
                double *rw = row(k);
                if (rw[RowF::ON] == 0.0) continue;
                double gq = 0.0, dvg = 0.0;
#pragma unroll
                for (int i = 0; i < NX; ++i) {
                    const double dv = upd ? w.DV[k * NZ + NU + i] : 0.0;
                    gq += rw[RowF::GC + i] * (w.DZ[k * NZ + NU + i] + as * dv);
                    dvg += rw[RowF::GC + i] * dv;
                }
                double lam[4], t[4], sg[2] = {rw[RowF::SIG], rw[RowF::SIG + 1]};
#pragma unroll
                for (int sd = 0; sd < 4; ++sd) lam[sd] = rw[RowF::LQ + sd], t[sd] = rw[RowF::TQ + sd];
                if (upd) {
                    // the steps of the accepted solve, recomputed as con_pass formed them
                    const double dsl = rw[RowF::DSIG], dsu = rw[RowF::DSIG + 1];
                    const double dtt[4] = {dvg + dsl - rw[RowF::RD], -dvg + dsu - rw[RowF::RD + 1], dsl - rw[RowF::RD + 2],
                                           dsu - rw[RowF::RD + 3]};
#pragma unroll
                    for (int sd = 0; sd < 4; ++sd) {
                        if (sd < nsd) {
                            double rm = lam[sd] * t[sd];
                            if (umode == 1) rm += rw[RowF::PR + sd] - usm;
                            if (umode == 2) rm -= usm;
                            const double dl = -(rm + lam[sd] * dtt[sd]) * VB_RCP(t[sd]);
                            lam[sd] = fmax(lam[sd] + as * dl, o.qp_lam_min);
                            t[sd] = fmax(t[sd] + as * dtt[sd], o.qp_t_min);
                        }
                    }
                    sg[0] += as * dsl, sg[1] += as * dsu;
                }
                const double rd[4] = {rw[RowF::LGD] - gq - sg[0] + t[0], gq - rw[RowF::UGD] - sg[1] + t[1], -sg[0] + t[2],
                                      -sg[1] + t[3]};
                double wv[4] = {0, 0, 0, 0}, beta[4] = {0, 0, 0, 0};
#pragma unroll
                for (int sd = 0; sd < 4; ++sd) {
                    if (sd < nsd) {
                        const double m_ = lam[sd] * t[sd], it = VB_RCP(t[sd]);
                        wv[sd] = lam[sd] * it, beta[sd] = (m_ - lam[sd] * rd[sd]) * it;
                        nb |= (rd[sd] != rd[sd]) | (m_ != m_);
                        vd = fmax(vd, fabs(rd[sd])), vm = fmax(vm, fabs(m_)), mu += m_;
                        rw[RowF::LQ + sd] = lam[sd], rw[RowF::TQ + sd] = t[sd], rw[RowF::RD + sd] = rd[sd];
                        lmx = fmax(lmx, lam[sd]);
                    }
                }
                double a1 = 1.0, a2 = 1.0, c3 = 0.0, c4 = 0.0, rsl = 0.0, rsu = 0.0;
                if (g->soft) {
                    rw[RowF::SIG] = sg[0], rw[RowF::SIG + 1] = sg[1];
                    rsl = rw[RowF::ZS] * sg[0] + rw[RowF::ZS + 2] - lam[0] - lam[2];
                    rsu = rw[RowF::ZS + 1] * sg[1] + rw[RowF::ZS + 3] - lam[1] - lam[3];
                    nb |= (rsl != rsl) | (rsu != rsu);
                    vg = fmax(vg, fmax(fabs(rsl), fabs(rsu)));
                    const double iDl = 1.0 / (rw[RowF::ZS] + wv[0] + wv[2]), iDu = 1.0 / (rw[RowF::ZS + 1] + wv[1] + wv[3]);
                    a1 = (rw[RowF::ZS] + wv[2]) * iDl, c3 = wv[0] * iDl;
                    a2 = (rw[RowF::ZS + 1] + wv[3]) * iDu, c4 = wv[1] * iDu;
                }
                rw[RowF::WEFF] = wv[0] * a1 + wv[1] * a2;
                rw[RowF::BARG] = (a1 * beta[0] - c3 * (rsl + beta[2])) - (a2 * beta[1] - c4 * (rsu + beta[3]));
            }
            L(a_g) = vg, L(a_d) = vd, L(a_m) = vm, L(a_mu) = mu, L(bad) = nb, L(a_lm) = lmx;
            END_LANES
        }
        FOR_LANES
        double vg = 0, vd = 0, vm = 0, mu = 0;
        int nb = 0;
        if constexpr (ROWS) vg = L(a_g), vd = L(a_d), vm = L(a_m), mu = L(a_mu), nb = L(bad);
#pragma unroll FLAT_UNROLL
        for (int idx = lane; idx < (N + 1) * NZ; idx += 32) {
            int k = idx / NZ, i = idx - k * NZ;
            {
                const int nidx = idx + VB_PF_DIST, nk = nidx / NZ, nc = CI(nk, nidx - nk * NZ, 0);
                constexpr int NZ_ = VB_PAIR_LAYOUT ? 1 : NZ;  // distance of the upper entry
                (void)NZ_;
                if (VB_PF_RES && nidx < (N + 1) * NZ) {
                    VB_PREFETCH(w.DZ + nidx), VB_PREFETCH(w.Z + nidx);
                    VB_PREFETCH(rec(nk) + R::BAT + (nidx - nk * NZ) * NX), VB_PREFETCH(w.PIQ + nk * NX);
                    VB_PREFETCH(w.LAMQ + nc), VB_PREFETCH(w.LAMQ + nc + NZ_);
                    VB_PREFETCH(w.TQ + nc), VB_PREFETCH(w.TQ + nc + NZ_);
                    if (upd) {
                        VB_PREFETCH(w.DV + nidx), VB_PREFETCH(w.DLAM + nc), VB_PREFETCH(w.DLAM + nc + NZ_);
                        VB_PREFETCH(w.DT + nc), VB_PREFETCH(w.DT + nc + NZ_);
                    }
                }
            }
            double v = w.DZ[idx];
            if (upd) {
                v += as * w.DV[idx];
                w.DZ[idx] = v;
            }
            double hh = cost_h(k, i) + o.qp_reg_prim, bar = 0.0;
            double r = (hh - o.qp_reg_prim) * v + cost_g(k, i, w.Z[idx]);
            double *rk = rec(k);
            if (k < N) {
                const double *col = rk + R::BAT + i * NX, *pi = w.PIQ + k * NX;
#if VB_VEC_RES && defined(__CUDA_ARCH__)
                // column i of [B A] and pi_k are 16-byte aligned runs of nx doubles: 128-bit loads
                static_assert(Work<NQ>::O_PIQ % 2 == 0 && R::BAT % 2 == 0 && R::SIZE % 2 == 0 && NX % 2 == 0, "alignment");
                const double2 *c2 = reinterpret_cast<const double2 *>(col), *p2 = reinterpret_cast<const double2 *>(pi);
#pragma unroll
                for (int m = 0; m < NX / 2; ++m) {
                    const double2 cv = c2[m], pv = p2[m];
                    r += cv.x * pv.x;
                    r += cv.y * pv.y;
                }
#else
#pragma unroll
                for (int m = 0; m < NX; ++m) r += col[m] * pi[m];
#endif
            }
            if (k > 0 && i >= NU) r -= w.PIQ[(k - 1) * NX + i - NU];
            const double *rw_ = nullptr;
            if constexpr (ROWS) {
                if (i >= NU) {
                    const double *rw = row(k);
                    if (rw[RowF::ON] != 0.0) {
                        rw_ = rw;
                        r += rw[RowF::GC + i - NU] * (rw[RowF::LQ + 1] - rw[RowF::LQ]);
                    }
                }
            }
            const int c = CI(k, i, 0), cu = CI(k, i, 1);
            if (active(k, i)) {
                int sc = sclass(k);
                double z = w.Z[idx];
#if VB_PAIR_LAYOUT && defined(__CUDA_ARCH__)
                const double2 l2 = *reinterpret_cast<const double2 *>(w.LAMQ + c), t2 = *reinterpret_cast<const double2 *>(w.TQ + c);
                double ll = l2.x, lu = l2.y, tl = t2.x, tu = t2.y;
                if (upd) {
#if VB_RECOMPUTE_STEP
                    // the slack / multiplier steps of the accepted solve, recomputed exactly as con_pass formed them
                    // (from the primal step, the old bound residuals and second-order products): DT / DLAM never
                    // travel through memory
                    const double2 rd2 = *reinterpret_cast<const double2 *>(w.RD + c);
                    double2 rm2 = make_double2(0.0, 0.0);
                    if (umode == 1) rm2 = *reinterpret_cast<const double2 *>(w.RM + c);
                    const double dvv = w.DV[idx];
                    double rml = ll * tl, rmu = lu * tu;
                    if (umode == 1) rml += rm2.x - usm, rmu += rm2.y - usm;
                    if (umode == 2) rml -= usm, rmu -= usm;
                    const double dtl = dvv - rd2.x, dtu = -dvv - rd2.y;
                    const double dll = -(rml + ll * dtl) * VB_RCP(tl), dlu = -(rmu + lu * dtu) * VB_RCP(tu);
                    ll = fmax(ll + as * dll, o.qp_lam_min), lu = fmax(lu + as * dlu, o.qp_lam_min);
                    tl = fmax(tl + as * dtl, o.qp_t_min), tu = fmax(tu + as * dtu, o.qp_t_min);
#else
                    const double2 dl2 = *reinterpret_cast<const double2 *>(w.DLAM + c), dt2 = *reinterpret_cast<const double2 *>(w.DT + c);
                    ll = fmax(ll + as * dl2.x, o.qp_lam_min), lu = fmax(lu + as * dl2.y, o.qp_lam_min);
                    tl = fmax(tl + as * dt2.x, o.qp_t_min), tu = fmax(tu + as * dt2.y, o.qp_t_min);
#endif
                    *reinterpret_cast<double2 *>(w.LAMQ + c) = make_double2(ll, lu);
                    *reinterpret_cast<double2 *>(w.TQ + c) = make_double2(tl, tu);
                }
#else
                double ll = w.LAMQ[c], lu = w.LAMQ[cu], tl = w.TQ[c], tu = w.TQ[cu];
                if (upd) {
#if VB_RECOMPUTE_STEP
                    const double dvv = w.DV[idx];
                    double rml = ll * tl, rmu = lu * tu;
                    if (umode == 1) rml += w.RM[c] - usm, rmu += w.RM[cu] - usm;
                    if (umode == 2) rml -= usm, rmu -= usm;
                    const double dtl = dvv - w.RD[c], dtu = -dvv - w.RD[cu];
                    const double dll = -(rml + ll * dtl) * VB_RCP(tl), dlu = -(rmu + lu * dtu) * VB_RCP(tu);
                    ll = fmax(ll + as * dll, o.qp_lam_min), lu = fmax(lu + as * dlu, o.qp_lam_min);
                    tl = fmax(tl + as * dtl, o.qp_t_min), tu = fmax(tu + as * dtu, o.qp_t_min);
#else
                    ll = fmax(ll + as * w.DLAM[c], o.qp_lam_min), lu = fmax(lu + as * w.DLAM[cu], o.qp_lam_min);
                    tl = fmax(tl + as * w.DT[c], o.qp_t_min), tu = fmax(tu + as * w.DT[cu], o.qp_t_min);
#endif
                    w.LAMQ[c] = ll, w.LAMQ[cu] = lu, w.TQ[c] = tl, w.TQ[cu] = tu;
                }
#endif
                r += lu - ll;
                double dl = (s.lb[sc][i] - z) - v + tl, du = v - (s.ub[sc][i] - z) + tu;
                double ml = ll * tl, mu_ = lu * tu;
                w.RD[c] = dl, w.RD[cu] = du;
                if (VB_STORE_RMB) w.RMB[c] = ml, w.RMB[cu] = mu_;
                nb |= (dl != dl) | (du != du) | (ml != ml) | (mu_ != mu_);
                vd = fmax(vd, fmax(fabs(dl), fabs(du)));
                vm = fmax(vm, fmax(fabs(ml), fabs(mu_)));
                mu += ml + mu_;
                double itl = VB_RCP(tl), itu = VB_RCP(tu);
                hh += ll * itl + lu * itu;
                bar = (ml - ll * dl) * itl - (mu_ - lu * du) * itu;
            } else {
                w.RD[c] = 0.0, w.RD[cu] = 0.0;
                if (VB_STORE_RMB) w.RMB[c] = 0.0, w.RMB[cu] = 0.0;
                if (k == N) r = 0.0;
            }
            if (VB_RG_ALL || k == 0) w.RG[idx] = r;  // only the stage-0 state part is read back (projection)
            if constexpr (ROWS) {
                if (rw_) bar += rw_[RowF::BARG] * rw_[RowF::GC + i - NU];
            }
            rk[R::HH + i] = hh;
            rk[R::RR + i] = r + bar;
            if (!(k == 0 && i >= NU)) {
                nb |= (r != r);
                vg = fmax(vg, fabs(r));
            }
        }
        L(a_g) = vg, L(a_d) = vd, L(a_m) = vm, L(a_mu) = mu, L(bad) = nb;
        END_LANES  // the loop below reads DZ entries that other lanes may just have updated
        FOR_LANES
        double vb = 0;
        int nb = 0;
        for (int idx = lane; idx < N * NX; idx += 32) {
            int k = idx / NX, i = idx - k * NX;
            double a = w.BD[idx] - w.DZ[(k + 1) * NZ + NU + i];
            const double *dz = w.DZ + k * NZ, *bat = rec(k) + R::BAT + i;
#pragma unroll
            for (int j = 0; j < NZ; ++j) a += bat[j * NX] * dz[j];
            rec(k)[R::RB + i] = a;
            nb |= (a != a);
            vb = fmax(vb, fabs(a));
        }
        L(a_b) = vb, L(bad) |= nb;
        END_LANES
        FOR_LANES
        if (lane < NX) {
            s.va[lane] = w.RG[NU + lane];
            s.vb[lane] = w.Z[NU + lane] + w.DZ[NU + lane];
            s.eN[lane] = ((s.fixedN >> lane) & 1)
                             ? w.Z[N * NZ + NU + lane] + w.DZ[N * NZ + NU + lane] - s.cN[lane]
                             : 0.0;
        }
        END_LANES
        proj0(s.va);
        eq0_violation(s.vb, s.e0);
        // the stage-0 pieces join the lane-wise maxima before the reductions (a max is order independent)
        FOR_LANES
        if (lane < NX) {
            const double va = s.va[lane];
            L(bad) |= (va != va);
            L(a_g) = fmax(L(a_g), fabs(va));
            L(a_b) = fmax(L(a_b), fmax(fabs(s.e0[lane]), fabs(s.eN[lane])));
        }
        END_LANES
        ng = WARP_MAX(a_g), nb_ = WARP_MAX(a_b), nd = WARP_MAX(a_d), nm = WARP_MAX(a_m);
        if constexpr (ROWS) {
            // Stationarity is tested RELATIVE to the largest row multiplier: with a strongly active row (multiplier ~1e4,
            // or Zl sl ~ 1e6 ... 1e12 for a penalised slack) the residual is a difference of numbers of that size, and the
            // barrier weight lam / t of the row conditions the Newton systems ~1e17 near the solution -- an absolute
            // 1e-8 is then below the rounding floor of FP64 and an IPM asked for it wanders off again after reaching it.
            ng /= fmax(1.0, WARP_MAX(a_lm));
        }
        double mu = WARP_SUM(a_mu);
        nan = WARP_ANY(bad);
        FOR_LANES
        if (lane < NX) {
            double rv = s.va[lane], rraw = w.RG[NU + lane];
            w.RG[NU + lane] = rv;
            rec(0)[R::RR + NU + lane] += rv - rraw;  // the stage-0 state gradient is the projected one
        }
        END_LANES
        PROXY_FENCE();  // the sweeps read the records through TMA
        return s.nact ? mu / (2.0 * s.nact) : 0.0;
    }

    // gradient of the Newton system in the three solve modes: 0 predictor, 1 corrector
    // (centering + second-order term), 2 centering only
    VB_DEV double rhs_of(const double *r, int i, int mode, double sm) const {
        if (mode == 0) return r[R::RR + i];
        return (mode == 1 ? r[R::Q1 + i] : r[R::RR + i]) - sm * r[R::Q2 + i];
    }

    // ---------------------------------------------------------------- Riccati: backward sweep
    // mode 0: factorise + solve the predictor; mode 1/2: re-use the factors with the corrector /
    // centering-only gradient.  Stage records arrive through the TMA ring (RING_DEPTH stages ahead).
    // Ends with the stage-0 step in s.dx[0].  Returns false on a singular terminal block.
    VB_DEV bool backward(int mode, double sm) {
        const int N = s.N;
        const bool factor = mode == 0;
        {
            const double *rN = rec(N);
            FOR_LANES
            if (lane < NX) {
                bool fx = (s.fixedN >> lane) & 1;
                double hh = rN[R::HH + NU + lane], rr = rhs_of(rN, NU + lane, mode, sm);
                s.hhN[lane] = fx ? 0.0 : hh;
                s.rN[lane] = fx ? 0.0 : rr;
                s.pvec[lane] = fx ? 0.0 : rr;
                if (factor) {
#pragma unroll
                    for (int j = 0; j < NX; ++j) {
                        double pv = (j == lane && !fx) ? hh : 0.0;
                        if constexpr (ROWS) {  // barrier Hessian of the terminal row
                            const double *rw = row(N);
                            if (rw[RowF::ON] != 0.0) pv += rw[RowF::WEFF] * rw[RowF::GC + lane] * rw[RowF::GC + j];
                        }
                        s.P[lane][j] = pv;
                    }
                }
            }
            END_LANES
        }
        // only the window of the record this sweep reads travels through the ring
        const int w0 = factor ? R::W_FAC0 : R::W_SOL0;
        const unsigned wbytes = (factor ? R::W_FAC1 - R::W_FAC0 : R::W_SOL1 - R::W_SOL0) * 8u;
        for (int j = 0; j < RING_DEPTH && j < N; ++j)
            RING_FETCH(ring, j, s.ring[j] + w0, rec(N - 1 - j) + w0, wbytes);
        bool ok = true;
        for (int it = 0, slot = 0; it < N; ++it, slot = slot + 1 == RING_DEPTH ? 0 : slot + 1) {
            const int k = N - 1 - it;
            const bool last = (it == 0) && s.termfix;
            RING_WAIT(ring, slot);
            const double *r = s.ring[slot];
            double *gk = rec(k), *pp = w.PP + (size_t)k * Work<NQ>::PPS;
            if (factor) {
                // B: P+ [B A]  and  tb = P+ beta
                FOR_LANES
                for (int idx = lane; idx < NZ * NX; idx += 32) {
                    int j = idx / NX, i = idx - j * NX;
                    s.PBAT[j][i] = dotv<NX>(&s.P[i][0], r + R::BAT + j * NX);
                }
                if (lane < NX) {
                    s.tb[lane] = dotv<NX>(&s.P[lane][0], r + R::RB);
                }
                END_LANES
            }
            // C: M = H + [B A]' P+ [B A]  and  m = r + [B A]'(P+ beta) + [B A]' p+
            FOR_LANES
            if (factor) {
                for (int idx = lane; idx < TRI; idx += 32) {
                    // (row, column) of the idx-th entry of the row-major lower triangle, from the table
                    const int ab = s.tri[idx], a_ = ab >> 4, b_ = ab & 15;
                    double a = dotv<NX>(r + R::BAT + a_ * NX, &s.PBAT[b_][0], (a_ == b_) ? r[R::HH + a_] : 0.0);
                    if constexpr (ROWS) {  // barrier Hessian of this stage's row (rank one in the states)
                        if (b_ >= NU) {
                            const double *rw = row(k);
                            if (rw[RowF::ON] != 0.0) a += rw[RowF::WEFF] * rw[RowF::GC + a_ - NU] * rw[RowF::GC + b_ - NU];
                        }
                    }
                    s.M[a_][b_] = a, s.M[b_][a_] = a;
                }
            } else if (last) {
                for (int idx = lane; idx < NZ * NZ; idx += 32) (&s.M[0][0])[idx] = w.MF[idx];
            }
            if (lane < NZ) {
                double mb;
                if (factor) {
                    mb = dotv<NX>(r + R::BAT + lane * NX, s.tb);
                    gk[R::MB + lane] = mb;
                } else {
                    mb = r[R::MB + lane];
                }
                s.m[lane] = dotv<NX>(r + R::BAT + lane * NX, s.pvec, rhs_of(r, lane, mode, sm) + mb);
            }
            END_LANES
            // D: eliminate the controls
            if (!last) {
                // nu x nu Cholesky, redundantly in every lane (strict lower part + inverse diagonal)
                double Lu[NU][NU], di[NU], y[NU];
                if (factor) {
#pragma unroll
                    for (int j = 0; j < NU; ++j) {
                        double d = s.M[j][j];
#pragma unroll
                        for (int c = 0; c < j; ++c) d -= Lu[j][c] * Lu[j][c];
                        di[j] = d > 0.0 ? VB_RSQRT(d) : 0.0;
#pragma unroll
                        for (int i = j + 1; i < NU; ++i) {
                            double a = s.M[i][j];
#pragma unroll
                            for (int c = 0; c < j; ++c) a -= Lu[i][c] * Lu[j][c];
                            Lu[i][j] = a * di[j];
                        }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < NU; ++i) {
                        di[i] = r[R::LUU + i * NU + i];
#pragma unroll
                        for (int c = 0; c < i; ++c) Lu[i][c] = r[R::LUU + i * NU + c];
                    }
                }
#pragma unroll
                for (int c = 0; c < NU; ++c) {
                    double a = s.m[c];
#pragma unroll
                    for (int c2 = 0; c2 < c; ++c2) a -= Lu[c][c2] * y[c2];
                    y[c] = a * di[c];
                }
                FOR_LANES
                if (factor) {
                    // P is symmetric: the nx (nx + 1) / 2 entries of its lower triangle fit one round of lanes,
                    // each entry is stored twice (the mirrored value is bit-identical: M is stored symmetric)
                    for (int idx = lane; idx < NX * (NX + 1) / 2; idx += 32) {
                        const int ab = s.tri[idx], i = ab >> 4, j = ab & 15;
                        double li[NU], lj[NU];
#pragma unroll
                        for (int c = 0; c < NU; ++c) {
                            double a = s.M[NU + i][c], b = s.M[NU + j][c];
#pragma unroll
                            for (int c2 = 0; c2 < c; ++c2) a -= li[c2] * Lu[c][c2], b -= lj[c2] * Lu[c][c2];
                            li[c] = a * di[c], lj[c] = b * di[c];
                        }
                        double a = s.M[NU + i][NU + j];
#pragma unroll
                        for (int c = 0; c < NU; ++c) a -= li[c] * lj[c];
                        s.P[i][j] = a, s.P[j][i] = a;
                        pp[i * NX + j] = a, pp[j * NX + i] = a;
                        if (j == 0) {
#pragma unroll
                            for (int c = 0; c < NU; ++c) gk[R::LXU + i * NU + c] = li[c];
                        }
                    }
                    if (lane == 0) {
#pragma unroll
                        for (int i = 0; i < NU; ++i)
#pragma unroll
                            for (int c = 0; c < NU; ++c)
                                gk[R::LUU + i * NU + c] = (c == i) ? di[i] : (c < i ? Lu[i][c] : 0.0);
                    }
                }
                if (lane < NX) {
                    double li[NU];
                    if (factor) {
#pragma unroll
                        for (int c = 0; c < NU; ++c) {
                            double a = s.M[NU + lane][c];
#pragma unroll
                            for (int c2 = 0; c2 < c; ++c2) a -= li[c2] * Lu[c][c2];
                            li[c] = a * di[c];
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < NU; ++c) li[c] = r[R::LXU + lane * NU + c];
                    }
                    double p = s.m[NU + lane];
#pragma unroll
                    for (int c = 0; c < NU; ++c) p -= li[c] * y[c];
                    s.pvec[lane] = p;
                    pp[NX * NX + lane] = p;
                }
                if (lane < NU) {
                    double yv = 0.0;
#pragma unroll
                    for (int c = 0; c < NU; ++c)
                        if (c == lane) yv = y[c];
                    gk[R::YV + lane] = yv;
                }
                END_LANES
            } else {
                // terminal equalities: G du + Gx dx + beta_v = -eN on the velocity rows determines
                // du = K dx + k0; the value function is the restriction of M to that manifold.
                double Gi[NU][NU], k0[NU];
                if (factor) {
                    double G[NU][NU];
#pragma unroll
                    for (int a = 0; a < NU; ++a)
#pragma unroll
                        for (int b = 0; b < NU; ++b) G[a][b] = r[R::BAT + b * NX + NQ + a];
                    ok = inverse_small(G, Gi) && ok;
                } else {
#pragma unroll
                    for (int a = 0; a < NU; ++a)
#pragma unroll
                        for (int b = 0; b < NU; ++b) Gi[a][b] = r[R::LUU + a * NU + b];
                }
#pragma unroll
                for (int a = 0; a < NU; ++a) {
                    double v = 0.0;
#pragma unroll
                    for (int b = 0; b < NU; ++b) v -= Gi[a][b] * (s.eN[NQ + b] + r[R::RB + NQ + b]);
                    k0[a] = v;
                }
                FOR_LANES
                for (int idx = lane; idx < NU * NX; idx += 32) {
                    int a = idx / NX, j = idx - a * NX;
                    double v = 0.0;
                    if (factor) {
#pragma unroll
                        for (int a2 = 0; a2 < NU; ++a2)
#pragma unroll
                            for (int b = 0; b < NU; ++b)
                                if (a2 == a) v -= Gi[a2][b] * r[R::BAT + (NU + j) * NX + NQ + b];
                    } else {
                        v = r[R::LXU + j * NU + a];
                    }
                    s.K[a][j] = v;
                }
                if (lane < NZ) {
                    double v = s.m[lane];
#pragma unroll
                    for (int a = 0; a < NU; ++a) v += s.M[lane][a] * k0[a];
                    s.tz[lane] = v;
                }
                END_LANES
                FOR_LANES
                if (factor) {
                    for (int idx = lane; idx < NX * NX; idx += 32) {
                        int i = idx / NX, j = idx - i * NX;
                        double a = s.M[NU + i][NU + j];
#pragma unroll
                        for (int c = 0; c < NU; ++c) {
                            double ki = s.K[c][i];
                            a += ki * s.M[c][NU + j] + s.M[NU + i][c] * s.K[c][j];
#pragma unroll
                            for (int c2 = 0; c2 < NU; ++c2) a += ki * s.M[c][c2] * s.K[c2][j];
                        }
                        pp[idx] = a;
                        s.P[i][j] = a;
                    }
                    for (int idx = lane; idx < NX * NU; idx += 32) {
                        int i = idx / NU, c = idx - i * NU;
                        gk[R::LXU + idx] = s.K[c][i];
                    }
                    for (int idx = lane; idx < NZ * NZ; idx += 32) w.MF[idx] = (&s.M[0][0])[idx];
                    if (lane == 0) {
#pragma unroll
                        for (int a = 0; a < NU; ++a)
#pragma unroll
                            for (int b = 0; b < NU; ++b) gk[R::LUU + a * NU + b] = Gi[a][b];
                    }
                }
                if (lane < NX) {
                    double p = s.tz[NU + lane];
#pragma unroll
                    for (int a = 0; a < NU; ++a) p += s.K[a][lane] * s.tz[a];
                    s.pvec[lane] = p;
                    pp[NX * NX + lane] = p;
                }
                if (lane < NZ) w.MF[NZ * NZ + lane] = s.m[lane];
                if (lane < NU) {
                    double kv = 0.0;
#pragma unroll
                    for (int c = 0; c < NU; ++c)
                        if (c == lane) kv = k0[c];
                    w.MF[NZ * NZ + NZ + lane] = kv;
                }
                END_LANES
            }
            if (it + RING_DEPTH < N) RING_FETCH(ring, slot, s.ring[slot] + w0, rec(k - RING_DEPTH) + w0, wbytes);
        }
        // stage 0:  dx0 = -e0 + Z0 dy,  (Z0'P0 Z0) dy = -Z0'(p0 - P0 e0)
        if (factor) {
            // T = P0 Z0 (into s.M as scratch), Pe = P0 e0
            FOR_LANES
            for (int idx = lane; idx < NX * NX; idx += 32) {
                int i = idx / NX, c = idx - i * NX;
                double a = 0.0;
#pragma unroll
                for (int j = 0; j < NX; ++j) a += s.P[i][j] * s.Z0[j][c];
                s.M[i][c] = a;
            }
            if (lane < NX) {
                double a = 0.0;
#pragma unroll
                for (int j = 0; j < NX; ++j) a += s.P[lane][j] * s.e0[j];
                s.Pe[lane] = a;
            }
            END_LANES
            FOR_LANES
            for (int idx = lane; idx < NX * NX; idx += 32) {
                int a_ = idx / NX, b_ = idx - a_ * NX;
                double a = 0.0;
#pragma unroll
                for (int i = 0; i < NX; ++i) a += s.Z0[i][a_] * s.M[i][b_];
                s.Lz[a_][b_] = a;
            }
            END_LANES
            // right-looking Cholesky of the (zero padded) reduced Hessian, one column per step
#pragma unroll 1
            for (int j = 0; j < NX; ++j) {
                double d = s.Lz[j][j];
                double di = d > 0.0 ? VB_RSQRT(d) : 0.0;
                UNIFORM_SYNC();
                FOR_LANES
                if (lane > j && lane < NX) s.Lz[lane][j] *= di;
                if (lane == j) s.dzi[j] = di;
                END_LANES
                FOR_LANES
                for (int idx = lane; idx < NX * NX; idx += 32) {
                    int i = idx / NX, c = idx - i * NX;
                    if (c > j && i >= c) s.Lz[i][c] -= s.Lz[i][j] * s.Lz[c][j];
                }
                END_LANES
            }
        }
        {
            double y[NX], dy[NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                double a = 0.0;
#pragma unroll
                for (int m = 0; m < NX; ++m) a -= s.Z0[m][i] * (s.pvec[m] - s.Pe[m]);
#pragma unroll
                for (int j = 0; j < i; ++j) a -= s.Lz[i][j] * y[j];
                y[i] = a * s.dzi[i];
            }
#pragma unroll
            for (int i = NX - 1; i >= 0; --i) {
                double a = y[i];
#pragma unroll
                for (int j = i + 1; j < NX; ++j) a -= s.Lz[j][i] * dy[j];
                dy[i] = a * s.dzi[i];
            }
            FOR_LANES
            if (lane < NX) {
                double a = -s.e0[lane];
#pragma unroll
                for (int c = 0; c < NX; ++c) a += s.Z0[lane][c] * dy[c];
                s.dx[0][lane] = a;
            }
            END_LANES
        }
        PROXY_FENCE();  // the forward sweep re-reads the records through TMA
        return ok;
    }

    // closed-form inverse of the nu x nu terminal block (uniform)
    VB_DEV bool inverse_small(const double (&G)[NU][NU], double (&Gi)[NU][NU]) const {
        if constexpr (NU == 1) {
            Gi[0][0] = 1.0 / G[0][0];
            return G[0][0] != 0.0 && G[0][0] == G[0][0];
        } else if constexpr (NU == 2) {
            double det = G[0][0] * G[1][1] - G[0][1] * G[1][0];
            double r = 1.0 / det;
            Gi[0][0] = G[1][1] * r, Gi[0][1] = -G[0][1] * r;
            Gi[1][0] = -G[1][0] * r, Gi[1][1] = G[0][0] * r;
            return det != 0.0 && det == det;
        } else {
            double c00 = G[1][1] * G[2][2] - G[1][2] * G[2][1];
            double c01 = G[1][2] * G[2][0] - G[1][0] * G[2][2];
            double c02 = G[1][0] * G[2][1] - G[1][1] * G[2][0];
            double det = G[0][0] * c00 + G[0][1] * c01 + G[0][2] * c02;
            double r = 1.0 / det;
            Gi[0][0] = c00 * r;
            Gi[1][0] = c01 * r;
            Gi[2][0] = c02 * r;
            Gi[0][1] = (G[0][2] * G[2][1] - G[0][1] * G[2][2]) * r;
            Gi[1][1] = (G[0][0] * G[2][2] - G[0][2] * G[2][0]) * r;
            Gi[2][1] = (G[0][1] * G[2][0] - G[0][0] * G[2][1]) * r;
            Gi[0][2] = (G[0][1] * G[1][2] - G[0][2] * G[1][1]) * r;
            Gi[1][2] = (G[0][2] * G[1][0] - G[0][0] * G[1][2]) * r;
            Gi[2][2] = (G[0][0] * G[1][1] - G[0][1] * G[1][0]) * r;
            return det != 0.0 && det == det;
        }
    }

    // ---------------------------------------------------------------- Riccati: forward sweep
    // DV from the factors: du_k = -Luu^-T (y_k + Lxu' dx_k), dx_{k+1} = [B A] dz_k + beta_k.  One
    // lane region per stage: every lane forms du redundantly in registers, lanes < nx form dx+.
    VB_DEV void forward() {
        const int N = s.N;
        constexpr int w0 = R::W_FWD0;
        constexpr unsigned wbytes = (R::W_FWD1 - R::W_FWD0) * 8u;
        for (int j = 0; j < RING_DEPTH && j < N; ++j) RING_FETCH(ring, j, s.ring[j] + w0, rec(j) + w0, wbytes);
        for (int k = 0, slot = 0; k < N; ++k, slot = slot + 1 == RING_DEPTH ? 0 : slot + 1) {
            const int cur = k & 1;
            const bool last = (k == N - 1) && s.termfix;
            RING_WAIT(ring, slot);
            const double *r = s.ring[slot];
            double du[NU];
            if (!last) {
                double t[NU];
#pragma unroll
                for (int c = 0; c < NU; ++c) {
                    t[c] = dot2<NX>(r + R::LXU + c, NU, s.dx[cur], 1, r[R::YV + c]);
                }
#pragma unroll
                for (int c = NU - 1; c >= 0; --c) {
                    double a = -t[c];
#pragma unroll
                    for (int c2 = c + 1; c2 < NU; ++c2) a -= r[R::LUU + c2 * NU + c] * du[c2];
                    du[c] = a * r[R::LUU + c * NU + c];
                }
            } else {
#pragma unroll
                for (int a_ = 0; a_ < NU; ++a_) {
                    double a = w.MF[NZ * NZ + NZ + a_];
#pragma unroll
                    for (int j = 0; j < NX; ++j) a += r[R::LXU + j * NU + a_] * s.dx[cur][j];
                    du[a_] = a;
                }
            }
            FOR_LANES
            if (lane < NX) {
                double a = dot2<NX>(r + R::BAT + NU * NX + lane, NX, s.dx[cur], 1, r[R::RB + lane]);
#pragma unroll
                for (int c = 0; c < NU; ++c) a += r[R::BAT + c * NX + lane] * du[c];
                s.dx[cur ^ 1][lane] = a;
                w.DV[k * NZ + NU + lane] = s.dx[cur][lane];
            } else if (lane < NZ) {
                double v = 0.0;
#pragma unroll
                for (int c = 0; c < NU; ++c)
                    if (c == lane - NX) v = du[c];
                w.DV[k * NZ + lane - NX] = v;
            }
            END_LANES
            if (last) {
                // multiplier of the eliminated velocity rows from the u-stationarity of the last stage
                double tu[NU], nuv[NU];
#pragma unroll
                for (int a_ = 0; a_ < NU; ++a_) {
                    double a = w.MF[NZ * NZ + a_];
#pragma unroll
                    for (int j = 0; j < NU; ++j) a += w.MF[a_ * NZ + j] * du[j];
#pragma unroll
                    for (int j = 0; j < NX; ++j) a += w.MF[a_ * NZ + NU + j] * s.dx[cur][j];
                    tu[a_] = a;
                }
#pragma unroll
                for (int b = 0; b < NU; ++b) {
                    double a = 0.0;
#pragma unroll
                    for (int a_ = 0; a_ < NU; ++a_) a -= r[R::LUU + a_ * NU + b] * tu[a_];
                    nuv[b] = a;
                }
                FOR_LANES
                if (lane < NU) {
                    double v = 0.0;
#pragma unroll
                    for (int c = 0; c < NU; ++c)
                        if (c == lane) v = nuv[c];
                    s.nuv[NQ + lane] = v;
                }
                END_LANES
            }
            if (k + RING_DEPTH < N) RING_FETCH(ring, slot, s.ring[slot] + w0, rec(k + RING_DEPTH) + w0, wbytes);
        }
        FOR_LANES
        if (lane < NZ) w.DV[N * NZ + lane] = lane < NU ? 0.0 : s.dx[N & 1][lane - NU];
        END_LANES
    }

    // ---------------------------------------------------------------- constraint pass
    // Steps of the slacks and multipliers from DV, the maximum step to the boundary, and the three
    // sums that give mu(alpha) = (S0 + alpha S1 + alpha^2 S2) / nc.  mode 0 (predictor) also stores
    // the second-order products and the corrector gradient pieces Q1, Q2; modes 1/2 store DT, DLAM.
    VB_DEV double con_pass(int mode, double sm, double &S0, double &S1, double &S2) {
        const int N = s.N;
        LV(double, amin);
        LV(double, a0);
        LV(double, a1);
        LV(double, a2);
        FOR_LANES
        double al = 1.0, s0 = 0, s1 = 0, s2 = 0;
        if constexpr (ROWS) {
            // the rows (lanes over the stages): slack steps from the eliminated slack equations, then the steps of the
            // four (two, if hard) one-sided constraints like those of a bound
            const int nsd = g->soft ? 4 : 2;
            for (int k = N - lane; k >= 0; k -= 32) {
                double *rw = row(k);
                if (rw[RowF::ON] == 0.0) continue;
                double dvg = 0.0;
#pragma unroll
                for (int i = 0; i < NX; ++i) dvg += rw[RowF::GC + i] * w.DV[k * NZ + NU + i];
                double lam[4], t[4], it[4] = {0, 0, 0, 0}, rm[4] = {0, 0, 0, 0}, wv[4] = {0, 0, 0, 0}, beta[4] = {0, 0, 0, 0};
#pragma unroll
                for (int sd = 0; sd < 4; ++sd) {
                    lam[sd] = rw[RowF::LQ + sd], t[sd] = rw[RowF::TQ + sd];
                    if (sd < nsd) {
                        rm[sd] = lam[sd] * t[sd];
                        if (mode == 1) rm[sd] += rw[RowF::PR + sd] - sm;
                        if (mode == 2) rm[sd] -= sm;
                        it[sd] = VB_RCP(t[sd]);
                        wv[sd] = lam[sd] * it[sd], beta[sd] = (rm[sd] - lam[sd] * rw[RowF::RD + sd]) * it[sd];
                    }
                }
                double dsl = 0.0, dsu = 0.0, a1 = 1.0, a2 = 1.0, c3 = 0.0, c4 = 0.0;
                if (g->soft) {
                    const double rsl = rw[RowF::ZS] * rw[RowF::SIG] + rw[RowF::ZS + 2] - lam[0] - lam[2];
                    const double rsu = rw[RowF::ZS + 1] * rw[RowF::SIG + 1] + rw[RowF::ZS + 3] - lam[1] - lam[3];
                    const double iDl = 1.0 / (rw[RowF::ZS] + wv[0] + wv[2]), iDu = 1.0 / (rw[RowF::ZS + 1] + wv[1] + wv[3]);
                    dsl = -(rsl + beta[0] + beta[2] + wv[0] * dvg) * iDl;
                    dsu = -(rsu + beta[1] + beta[3] - wv[1] * dvg) * iDu;
                    a1 = (rw[RowF::ZS] + wv[2]) * iDl, c3 = wv[0] * iDl;
                    a2 = (rw[RowF::ZS + 1] + wv[3]) * iDu, c4 = wv[1] * iDu;
                }
                rw[RowF::DSIG] = dsl, rw[RowF::DSIG + 1] = dsu;
                const double dtt[4] = {dvg + dsl - rw[RowF::RD], -dvg + dsu - rw[RowF::RD + 1], dsl - rw[RowF::RD + 2],
                                       dsu - rw[RowF::RD + 3]};
                const double cf[4] = {a1, -a2, -c3, c4};  // how the four second-order / centering terms enter the x gradient
                double q1g = 0.0, q2g = 0.0;
#pragma unroll
                for (int sd = 0; sd < 4; ++sd) {
                    if (sd < nsd) {
                        const double dl = -(rm[sd] + lam[sd] * dtt[sd]) * it[sd];
                        if (dtt[sd] < 0.0 && t[sd] + al * dtt[sd] < 0.0) al = fmin(al, VB_RATIO(t[sd], -dtt[sd]));
                        if (dl < 0.0 && lam[sd] + al * dl < 0.0) al = fmin(al, VB_RATIO(lam[sd], -dl));
                        s0 += lam[sd] * t[sd], s1 += lam[sd] * dtt[sd] + t[sd] * dl, s2 += dtt[sd] * dl;
                        if (mode == 0) {
                            const double pr = dtt[sd] * dl;
                            rw[RowF::PR + sd] = pr;
                            q1g += cf[sd] * pr * it[sd], q2g += cf[sd] * it[sd];
                        }
                    }
                }
                if (mode == 0) rw[RowF::Q1G] = q1g, rw[RowF::Q2G] = q2g;
            }
        }
#pragma unroll FLAT_UNROLL
        for (int idx = lane; idx < (N + 1) * NZ; idx += 32) {
            int k = idx / NZ, i = idx - k * NZ;
            double *rk = rec(k);
            double q1 = 0.0, q2 = 0.0;
            {
                const int nidx = idx + VB_PF_DIST, nk = nidx / NZ, nc = CI(nk, nidx - nk * NZ, 0);
                if (VB_PF_CON && nidx < (N + 1) * NZ) {
                    VB_PREFETCH(w.DV + nidx);
#if VB_PAIR_LAYOUT
                    VB_PREFETCH(w.LAMQ + nc), VB_PREFETCH(w.TQ + nc), VB_PREFETCH(w.RD + nc);  // pairs share a sector
                    if (mode == 1) VB_PREFETCH(w.RM + nc);
#else
                    VB_PREFETCH(w.LAMQ + nc), VB_PREFETCH(w.LAMQ + nc + NZ);
                    VB_PREFETCH(w.TQ + nc), VB_PREFETCH(w.TQ + nc + NZ);
                    if (VB_STORE_RMB) VB_PREFETCH(w.RMB + nc), VB_PREFETCH(w.RMB + nc + NZ);
                    VB_PREFETCH(w.RD + nc), VB_PREFETCH(w.RD + nc + NZ);
                    if (mode == 1) VB_PREFETCH(w.RM + nc), VB_PREFETCH(w.RM + nc + NZ);
#endif
                    if (mode == 0) VB_PREFETCH(rec(nk) + R::RR + (nidx - nk * NZ));
                }
            }
            if (active(k, i)) {
                double dvv = w.DV[idx];
#if VB_PAIR_LAYOUT && defined(__CUDA_ARCH__)
                // both sides of the component with one 128-bit load per array
                const int c0 = CI(k, i, 0);
                const double2 l2 = *reinterpret_cast<const double2 *>(w.LAMQ + c0);
                const double2 t2 = *reinterpret_cast<const double2 *>(w.TQ + c0);
                const double2 rd2 = *reinterpret_cast<const double2 *>(w.RD + c0);
                double2 rm2 = make_double2(0.0, 0.0);
                if (mode == 1) rm2 = *reinterpret_cast<const double2 *>(w.RM + c0);
#endif
#if VB_CON_ROLLED
#pragma unroll 1
#else
#pragma unroll
#endif
                for (int sd = 0; sd < 2; ++sd) {
                    const int c = CI(k, i, sd);
#if VB_PAIR_LAYOUT && defined(__CUDA_ARCH__)
                    double lam = sd ? l2.y : l2.x, t = sd ? t2.y : t2.x, rm = lam * t;
                    if (mode == 1) rm += (sd ? rm2.y : rm2.x) - sm;
                    if (mode == 2) rm -= sm;
                    double dtt = (sd ? -dvv : dvv) - (sd ? rd2.y : rd2.x);
#else
                    double lam = w.LAMQ[c], t = w.TQ[c], rm = VB_STORE_RMB ? w.RMB[c] : lam * t;
                    if (mode == 1) rm += w.RM[c] - sm;
                    if (mode == 2) rm -= sm;
                    double dtt = (sd ? -dvv : dvv) - w.RD[c];
#endif
                    double it = VB_RCP(t);
                    double dl = -(rm + lam * dtt) * it;
                    // ratio test; the division only when this constraint tightens the step
                    if (dtt < 0.0 && t + al * dtt < 0.0) al = fmin(al, VB_RATIO(t, -dtt));
                    if (dl < 0.0 && lam + al * dl < 0.0) al = fmin(al, VB_RATIO(lam, -dl));
                    s0 += lam * t, s1 += lam * dtt + t * dl, s2 += dtt * dl;
                    if (mode == 0) {
                        double pr = dtt * dl;
                        w.RM[c] = pr;
                        q1 += sd ? -pr * it : pr * it;
                        q2 += sd ? -it : it;
                    } else if (!VB_RECOMPUTE_STEP) {
                        w.DT[c] = dtt, w.DLAM[c] = dl;
                    }
                }
            }
            if (mode == 0) rk[R::Q1 + i] = rk[R::RR + i] + q1, rk[R::Q2 + i] = q2;
        }
        L(amin) = al, L(a0) = s0, L(a1) = s1, L(a2) = s2;
        END_LANES
        S0 = WARP_SUM(a0), S1 = WARP_SUM(a1), S2 = WARP_SUM(a2);
        double alpha = WARP_MIN(amin);
        if constexpr (ROWS) {
            if (mode == 0) {  // the rows' pieces of the corrector gradient act on x_k through the row gradient
                FOR_LANES
                for (int idx = lane; idx < (N + 1) * NX; idx += 32) {
                    const int k = idx / NX, i = idx - k * NX;
                    const double *rw = row(k);
                    if (rw[RowF::ON] != 0.0) {
                        rec(k)[R::Q1 + NU + i] += rw[RowF::GC + i] * rw[RowF::Q1G];
                        rec(k)[R::Q2 + NU + i] += rw[RowF::GC + i] * rw[RowF::Q2G];
                    }
                }
                END_LANES
            }
        }
        if (mode == 0) PROXY_FENCE();
        return alpha;
    }

    // ---------------------------------------------------------------- IPM
    // returns 0 success, 1 max iter, 2 min step, 3 NaN
    VB_DEV int ipm_solve(int &iters) {
        const int N = s.N;
        qp_init();
        double rg = 0, rb = 0, rd = 0, rm = 0, alpha = 1.0, mu = 0.0;
        bool nan = false, ok = true;
        const double nc = 2.0 * s.nact;
        int kk = 0;
        bool upd = false;
        double as_prev = 0.0, sm_prev = 0.0;
        int mode_prev = 1;
        for (;; ++kk) {
            mu = qp_residuals(rg, rb, rd, rm, nan, upd, as_prev, mode_prev, sm_prev);
            if (!(kk < o.qp_iter_max && alpha > o.qp_alpha_min && !nan &&
                  (rg > o.qp_tol_stat || rb > o.qp_tol_eq || rd > o.qp_tol_ineq || rm > o.qp_tol_comp)))
                break;
            // phase 0 predictor (res_m = lam*t), phase 1 centering + corrector
            // (res_m = lam*t + dt_aff*dlam_aff - sigma*mu), phase 2 centering only, taken when the
            // corrected step is much worse than the affine one (conditional predictor-corrector)
            double sm = 0.0, m_aff = 0.0;
#pragma unroll 1
            for (int ph = 0; ph < 3; ++ph) {
                double S0, S1, S2;
                ok = backward(ph, sm) && ok;
                if (!ok) break;
                forward();
                alpha = con_pass(ph, sm, S0, S1, S2);
                mode_prev = ph, sm_prev = sm;  // the solve whose step is applied (by the next residual pass)
                double m_a = (S0 + alpha * S1 + alpha * alpha * S2) / nc;
                if (ph == 0) {
                    m_aff = m_a;
                    double sigma = m_aff / mu;
                    sigma = sigma * sigma * sigma;
                    sm = sigma * mu;
                    if (sm < o.qp_tau_min) sm = o.qp_tau_min;
                } else if (ph == 1) {
                    if (!(m_a > 2.0 * m_aff)) break;
                }
            }
            if (!ok) break;
            double as = alpha;
            if (as < 1.0) as = as * ((1.0 - as) * 0.99 + as * 0.9999);
            // update: only the multipliers of the dynamics are stepped here, dpi_k = P_{k+1} dx_{k+1} + p_{k+1} from the
            // value functions; z, lam, t are stepped inside the next residual pass (qp_residuals(.., upd, as))
            upd = true, as_prev = as;
            FOR_LANES
#pragma unroll FLAT_UNROLL
            for (int idx = lane; idx < N * NX; idx += 32) {
                int k = idx / NX, mI = idx - k * NX;
                const double *dxn = w.DV + (size_t)(k + 1) * NZ + NU;
                if (VB_PF_UPD && idx + VB_PF_DIST < N * NX) {
                    const int nk = (idx + VB_PF_DIST) / NX, nm = idx + VB_PF_DIST - nk * NX;
                    VB_PREFETCH(w.PP + (size_t)(nk + 1) * Work<NQ>::PPS + nm * NX);
                    VB_PREFETCH(w.PP + (size_t)(nk + 1) * Work<NQ>::PPS + NX * NX + nm);
                    VB_PREFETCH(w.PIQ + idx + VB_PF_DIST);
                }
                double a;
                if (k + 1 < N) {
                    const double *pp = w.PP + (size_t)(k + 1) * Work<NQ>::PPS;
                    a = pp[NX * NX + mI];
#pragma unroll
                    for (int j = 0; j < NX; ++j) a += pp[mI * NX + j] * dxn[j];
                } else if ((s.fixedN >> mI) & 1) {
                    a = s.nuv[mI];
                } else {
                    a = s.hhN[mI] * dxn[mI] + s.rN[mI];
                    if constexpr (ROWS) {
                        const double *rw = row(N);
                        if (rw[RowF::ON] != 0.0) {
                            double gd = 0.0;
#pragma unroll
                            for (int j = 0; j < NX; ++j) gd += rw[RowF::GC + j] * dxn[j];
                            a += rw[RowF::WEFF] * rw[RowF::GC + mI] * gd;
                        }
                    }
                }
                w.PIQ[idx] += as * a;
            }
            END_LANES
        }
        iters = kk;
        // multipliers of the eliminated equalities from stationarity
        FOR_LANES
        if (lane < NX) {
            const int i = lane;
            double a = cost_h(0, NU + i) * w.DZ[NU + i] + cost_g(0, NU + i, w.Z[NU + i]);
            const double *col = rec(0) + R::BAT + (NU + i) * NX;
#pragma unroll
            for (int m = 0; m < NX; ++m) a += col[m] * w.PIQ[m];
            if (active(0, NU + i)) a += w.LAMQ[CI(0, NU + i, 1)] - w.LAMQ[CI(0, NU + i, 0)];
            if constexpr (ROWS) {
                const double *rw = row(0);
                if (rw[RowF::ON] != 0.0) a += rw[RowF::GC + i] * (rw[RowF::LQ + 1] - rw[RowF::LQ]);
            }
            s.va[i] = s.vb[i] = a;
            s.nuNq[i] = ((s.fixedN >> i) & 1)
                            ? w.PIQ[(N - 1) * NX + i] - cost_g(N, NU + i, w.Z[N * NZ + NU + i]) -
                                  cost_h(N, NU + i) * w.DZ[N * NZ + NU + i]
                            : 0.0;
        }
        END_LANES
        proj0(s.va);
        FOR_LANES
        if (lane < NX) s.nu0q[lane] = s.vb[lane] - s.va[lane];
        END_LANES
        if (!ok || nan || mu != mu) return 3;
        if (rg > o.qp_tol_stat || rb > o.qp_tol_eq || rd > o.qp_tol_ineq || rm > o.qp_tol_comp)
            return kk >= o.qp_iter_max ? 1 : 2;
        return 0;
    }

    // ---------------------------------------------------------------- merit function
    VB_DEV double total_cost(const double *Zs) {
        const int N = s.N;
        if (VBOCLIKE) {
            double c = s.wtdt;
#pragma unroll
            for (int i = 0; i < NQ; ++i) c += s.w[i] * Zs[NU + NQ + i];
            return c;
        }
        LV(double, acc);
        FOR_LANES
        double a = 0.0;
        for (int k = lane; k <= N; k += 32) {
            double q = 0.0;
            if constexpr (MPC) {
                // LINEAR_LS: 1/2 (y - y_ref)' W (y - y_ref), stage costs times the time step
                if (k < N) {
#pragma unroll
                    for (int i = 0; i < NZ; ++i) {
                        const double d = Zs[k * NZ + i] - g->yref[i];
                        q += g->Wz[i] * d * d;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < NX; ++i) {
                        const double d = Zs[k * NZ + NU + i] - g->yrefN[i];
                        q += g->WzN[i] * d * d;
                    }
                }
                q *= 0.5;
            } else {
#pragma unroll
                for (int i = 0; i < NQ; ++i) q += Zs[k * NZ + NU + NQ + i] * Zs[k * NZ + NU + NQ + i];
            }
            a += (k < N ? s.h : 1.0) * q;
        }
        L(acc) = a;
        END_LANES
        return WARP_SUM(acc);
    }

    VB_DEV double merit(const double *Zs) {
        const int N = s.N;
        LV(double, acc);
        FOR_LANES
        double a = 0.0;
        for (int k = lane; k < N; k += 32) {
            double xn[NX];
            rk4_step<NQ, double>(Zs + k * NZ + NU, Zs + k * NZ, s.h, xn);
#pragma unroll
            for (int i = 0; i < NX; ++i) a += w.WDYN[k * NX + i] * fabs(xn[i] - Zs[(k + 1) * NZ + NU + i]);
        }
        for (int idx = lane; idx < (N + 1) * NZ; idx += 32) {
            int k = idx / NZ, i = idx - k * NZ;
            if (active(k, i)) {
                int sc = sclass(k);
                double z = Zs[idx], fl = s.lb[sc][i] - z, fu = z - s.ub[sc][i];
                if (fl > 0) a += w.WB[CI(k, i, 0)] * fl;
                if (fu > 0) a += w.WB[CI(k, i, 1)] * fu;
            }
        }
        L(acc) = a;
        END_LANES
        double m = WARP_SUM(acc) + total_cost(Zs);
        if constexpr (ROWS) {
            // rows at the trial point: violation against the trial slacks (RowF::RD, set by line_search) and the
            // slacks' own cost; the trial slacks are convex combinations of non-negative values
            m += slack_cost(true);
#pragma unroll 1
            for (int k = row_first(); k <= row_last(); ++k) {
                const double h = row_value(Zs, k);
                const double *rw = row(k);
                const double fl = g->lh - h - rw[RowF::RD], fu = h - rw[RowF::RD + 1] - g->uh;
                if (fl > 0.0) m += rw[RowF::WGM] * fl;
                if (fu > 0.0) m += rw[RowF::WGM + 1] * fu;
            }
        }
        FOR_LANES
        if (lane < NX) s.vb[lane] = Zs[NU + lane];
        END_LANES
        eq0_violation(s.vb, s.vc);
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            m += s.w0[i] * fabs(s.vc[i]);
            if ((s.fixedN >> i) & 1) m += s.wN[i] * fabs(Zs[N * NZ + NU + i] - s.cN[i]);
        }
        UNIFORM_SYNC();
        return m;
    }

    VB_DEV double line_search(int sqp_iter, int &evals) {
        const int N = s.N;
        FOR_LANES
        for (int idx = lane; idx < N * NX; idx += 32) {
            double a = fabs(w.PIQ[idx]);
            w.WDYN[idx] = sqp_iter == 0 ? a : fmax(a, 0.5 * (w.WDYN[idx] + a));
        }
        for (int idx = lane; idx < (N + 1) * NC; idx += 32) {
            double a = fabs(w.LAMQ[idx]);
            w.WB[idx] = sqp_iter == 0 ? a : fmax(a, 0.5 * (w.WB[idx] + a));
        }
        if (lane < NX) {
            double a = fabs(s.nu0q[lane]), b = fabs(s.nuNq[lane]);
            s.w0[lane] = sqp_iter == 0 ? a : fmax(a, 0.5 * (s.w0[lane] + a));
            s.wN[lane] = sqp_iter == 0 ? b : fmax(b, 0.5 * (s.wN[lane] + b));
        }
        if constexpr (ROWS) {
            for (int idx = lane; idx < (N + 1) * 4; idx += 32) {
                double *rw = row(idx >> 2);
                const int sd = idx & 3;  // WGM[2] and WSM[2] are adjacent, like LQ[0..3]
                const double a = fabs(rw[RowF::LQ + sd]);
                rw[RowF::WGM + sd] = sqp_iter == 0 ? a : fmax(a, 0.5 * (rw[RowF::WGM + sd] + a));
            }
        }
        END_LANES
        // trial -1 evaluates the merit function at the current iterate (alpha = 0)
        double m0 = 0.0, alpha = 0.0;
#pragma unroll 1
        for (int trial = -1;; ++trial) {
            FOR_LANES
            for (int idx = lane; idx < (N + 1) * NZ; idx += 32) w.ZT[idx] = w.Z[idx] + alpha * w.DZ[idx];
            if constexpr (ROWS) {
                for (int idx = lane; idx < (N + 1) * 2; idx += 32) {  // trial slacks (the QP's SIG is the new slack VALUE)
                    double *rw = row(idx >> 1);
                    const int sd = idx & 1;
                    rw[RowF::RD + sd] = rw[RowF::SL + sd] + alpha * (rw[RowF::SIG + sd] - rw[RowF::SL + sd]);
                }
            }
            END_LANES
            double m1 = merit(w.ZT);
            if (trial < 0) {
                m0 = m1, alpha = 1.0;
            } else {
                ++evals;
                if (m1 < m0) break;
                if (alpha * o.alpha_reduction < o.alpha_min) break;  // the smallest step is taken anyway
                alpha *= o.alpha_reduction;
            }
        }
        return alpha;
    }

    // ---------------------------------------------------------------- SQP / RTI driver
    VB_DEV void solve(const Prob &pb, int mode) {
        load_problem(pb);
        const int N = s.N;
        vboc_stats st;
        st.status = VBOC_MAXITER, st.sqp_iter = 0, st.qp_iter = 0, st.ls_evals = 0, st.qp_status = 0;
        st.pad_ = 0;
        st.res_stat = st.res_eq = st.res_ineq = st.res_comp = 0.0;
        const int maxit = mode == VBOC_MODE_RTI ? 1 : o.max_iter;
        for (int it = 0;; ++it) {
            linearize();
            linearize_rows();
            bool finite = nlp_residuals(st.res_stat, st.res_eq, st.res_ineq, st.res_comp);
            if (mode == VBOC_MODE_SQP || it > 0) {
                if (!finite) {
                    st.status = VBOC_FAILURE;
                    break;
                }
                if (mode == VBOC_MODE_SQP && st.res_stat < o.tol_stat && st.res_eq < o.tol_eq &&
                    st.res_ineq < o.tol_ineq && st.res_comp < o.tol_comp) {
                    st.status = VBOC_SUCCESS;
                    break;
                }
            }
            if (it >= maxit) {
                st.status = mode == VBOC_MODE_RTI ? VBOC_SUCCESS : VBOC_MAXITER;
                break;
            }
            int qit = 0;
            int qs = ipm_solve(qit);
            st.qp_iter += qit, st.qp_status = qs, st.sqp_iter = it + 1;
            if (qs != 0 && qs != 1) {  // min step / NaN are fatal, max iter is tolerated
                st.status = VBOC_QP_FAILURE;
                break;
            }
            double alpha = 1.0;
            if (mode == VBOC_MODE_SQP && o.globalization) alpha = line_search(it, st.ls_evals);
            FOR_LANES
            for (int idx = lane; idx < (N + 1) * NZ; idx += 32) w.Z[idx] += alpha * w.DZ[idx];
            for (int idx = lane; idx < N * NX; idx += 32)
                w.PI[idx] = (1.0 - alpha) * w.PI[idx] + alpha * w.PIQ[idx];
            for (int idx = lane; idx < (N + 1) * NC; idx += 32)
                w.LAM[idx] = (1.0 - alpha) * w.LAM[idx] + alpha * w.LAMQ[idx];
            if constexpr (ROWS) {
                for (int idx = lane; idx < (N + 1) * 6; idx += 32) {
                    double *rw = row(idx / 6);
                    const int j = idx - (idx / 6) * 6;  // LAMG[2], LAMS[2] <- LQ[4]; SL[2] <- SIG[2]
                    const double qv = j < 4 ? rw[RowF::LQ + j] : rw[RowF::SIG + j - 4];
                    rw[RowF::LAMG + j] = (1.0 - alpha) * rw[RowF::LAMG + j] + alpha * qv;
                }
            }
            END_LANES
        }
        st.cost = total_cost(w.Z) + slack_cost(false);
        store_solution(pb, st);
    }
};

}  // namespace vboc
