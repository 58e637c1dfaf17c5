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
#include "warp_spmd.h"

namespace vboc {

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
};

template <int NQ>
struct Dim {
    static constexpr int NX = 2 * NQ, NU = NQ, NZ = 3 * NQ, NC = 2 * NZ;
    static constexpr int FS = NU * NU + NX * NU + NX * NX;  // Riccati factor record per stage
    static constexpr int MFS = NZ * NZ + NZ + NU;           // last-stage record
};

// Global-memory workspace of one warp slot, stage-major.
template <int NQ>
struct Work {
    using D = Dim<NQ>;
    double *Z, *PI, *LAM;            // NLP iterate: z_k = [u_k; x_k], dynamics and bound multipliers
    double *BAT, *BD;                // [B A]' per stage (nz x nx: column j of [B A] contiguous), gap
    double *DZ, *PIQ, *LAMQ, *TQ;    // QP iterate
    double *DV, *DPI, *DLAM, *DT;    // IPM step
    double *RG, *RB, *RD, *RM, *RMB; // IPM residuals (RMB = lam*t, RM = corrector rhs)
    double *FAC, *PV, *YV, *MF;      // Riccati factors and vectors
    double *WDYN, *WB, *ZT;          // merit weights, trial point
    static VB_HD size_t doubles(int Nmax) {
        size_t S = (size_t)Nmax + 1;
        return S * (size_t)(D::NZ + D::NX + D::NC + D::NZ * D::NX + D::NX + D::NZ + D::NX + D::NC +
                            D::NC + D::NZ + D::NX + D::NC + D::NC + D::NZ + D::NX + D::NC + D::NC +
                            D::NC + D::FS + D::NX + D::NU + D::NX + D::NC + D::NZ) +
               D::MFS;
    }
    VB_HD void carve(double *b, int Nmax) {
        size_t S = (size_t)Nmax + 1;
        auto take = [&](size_t per) {
            double *p = b;
            b += S * per;
            return p;
        };
        Z = take(D::NZ), PI = take(D::NX), LAM = take(D::NC);
        BAT = take(D::NZ * D::NX), BD = take(D::NX);
        DZ = take(D::NZ), PIQ = take(D::NX), LAMQ = take(D::NC), TQ = take(D::NC);
        DV = take(D::NZ), DPI = take(D::NX), DLAM = take(D::NC), DT = take(D::NC);
        RG = take(D::NZ), RB = take(D::NX), RD = take(D::NC), RM = take(D::NC), RMB = take(D::NC);
        FAC = take(D::FS), PV = take(D::NX), YV = take(D::NU);
        WDYN = take(D::NX), WB = take(D::NC), ZT = take(D::NZ);
        MF = b;
    }
};

// Shared-memory block of one warp.
template <int NQ>
struct Smem {
    using D = Dim<NQ>;
    // problem constants
    double lb[3][D::NZ], ub[3][D::NZ];  // stage classes 0, 1..N-1, N in z ordering
    double Z0[D::NX][D::NX];            // orthonormal basis of the stage-0 free subspace, zero padded
    double c0[D::NX], cN[D::NX], w[NQ];
    double h, wtdt;
    int N, fixed0, fixedN, termfix, nact;
    // Riccati staging
    double BAT[D::NZ][D::NX], P[D::NX][D::NX], PBAT[D::NZ][D::NX], M[D::NZ][D::NZ];
    double hh[D::NZ], rr[D::NZ], t2[D::NX], m[D::NZ], pvec[D::NX], beta[D::NX], dz[D::NZ], dxn[D::NX];
    double e0[D::NX], eN[D::NX], hhN[D::NX], rN[D::NX], Lz[D::NX][D::NX];
    // merit weights / multipliers of the eliminated equalities
    double w0[D::NX], wN[D::NX], nu0q[D::NX], nuNq[D::NX];
};

template <int NQ, int FAM>
struct WarpSolver {
    using D = Dim<NQ>;
    static constexpr int NX = D::NX, NU = D::NU, NZ = D::NZ, NC = D::NC, FS = D::FS;
    static constexpr int OFF_LXU = NU * NU, OFF_P = NU * NU + NX * NU;
    static constexpr int TRI = NZ * (NZ + 1) / 2;

    Smem<NQ> &s;
    Work<NQ> w;
    const vboc_opts &o;

    VB_DEV WarpSolver(Smem<NQ> &s_, const Work<NQ> &w_, const vboc_opts &o_) : s(s_), w(w_), o(o_) {}

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
        if (FAM == VBOC_FAMILY_VBOC) return (k == 0 && i >= NU + NQ) ? s.w[i - NU - NQ] : 0.0;
        return (i >= NU + NQ) ? 2.0 * (k < s.N ? s.h : 1.0) * zval : 0.0;
    }
    VB_DEV double cost_h(int k, int i) const {
        double hd = o.levenberg_marquardt;
        if (FAM == VBOC_FAMILY_AL && i >= NU + NQ) hd += 2.0 * (k < s.N ? s.h : 1.0);
        return hd;
    }
    // e = (I - Z0 Z0')(x - c0): violation of the stage-0 equalities (uniform)
    VB_DEV void eq0_violation(const double *x, double *e) const {
        double y[NX];
#pragma unroll
        for (int c = 0; c < NX; ++c) {
            double a = 0.0;
#pragma unroll
            for (int i = 0; i < NX; ++i) a += s.Z0[i][c] * (x[i] - s.c0[i]);
            y[c] = a;
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            double a = x[i] - s.c0[i];
#pragma unroll
            for (int c = 0; c < NX; ++c) a -= s.Z0[i][c] * y[c];
            e[i] = a;
        }
    }
    // v <- Z0 Z0' v (uniform)
    VB_DEV void proj0(double *v) const {
        double y[NX], out[NX];
#pragma unroll
        for (int c = 0; c < NX; ++c) {
            double a = 0.0;
#pragma unroll
            for (int i = 0; i < NX; ++i) a += s.Z0[i][c] * v[i];
            y[c] = a;
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            double a = 0.0;
#pragma unroll
            for (int c = 0; c < NX; ++c) a += s.Z0[i][c] * y[c];
            out[i] = a;
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) v[i] = out[i];
    }

    // ---------------------------------------------------------------- problem load / store
    VB_DEV void load_problem(const Prob &pb) {
        const int N = pb.N;
        FOR_LANES
        if (lane == 0) {
            s.N = N, s.h = pb.h;
            s.wtdt = (FAM == VBOC_FAMILY_VBOC) ? pb.wt * pb.h * N : 0.0;
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
            for (int i = 0; i < NQ; ++i) s.w[i] = (FAM == VBOC_FAMILY_VBOC) ? pb.p[i] : 0.0;
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
        for (int idx = lane; idx < N * NX; idx += 32) w.PI[idx] = 0.0;
        for (int idx = lane; idx < (N + 1) * NC; idx += 32) w.LAM[idx] = 0.0;
        END_LANES
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
        END_LANES
    }

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
                    double *col = w.BAT + ((size_t)k * NZ + dir) * NX;
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
                    const double *col = w.BAT + (size_t)idx * NX, *pi = w.PI + k * NX;
#pragma unroll
                    for (int m = 0; m < NX; ++m) r += col[m] * pi[m];
                }
                if (k > 0 && i >= NU) r -= w.PI[(k - 1) * NX + i - NU];
                if (active(k, i)) {
                    int sc = sclass(k);
                    double ll = w.LAM[k * NC + i], lu = w.LAM[k * NC + NZ + i];
                    double fl = s.lb[sc][i] - z, fu = z - s.ub[sc][i];
                    r += lu - ll;
                    vi = fmax(vi, fmax(fl, fu));
                    vc = fmax(vc, fmax(fabs(ll * fl), fabs(lu * fu)));
                } else if (k == N) {
                    r = 0.0;
                    vi = fmax(vi, fabs(z - s.cN[i - NU]));
                }
            }
            w.RG[idx] = r;  // scratch: the stage-0 state part is projected below
            if (!(k == 0 && i >= NU)) {
                nb |= (r != r);
                vs = fmax(vs, fabs(r));
            }
        }
        for (int idx = lane; idx < N * NX; idx += 32) {
            double v = w.BD[idx];
            nb |= (v != v);
            ve = fmax(ve, fabs(v));
        }
        L(a_s) = vs, L(a_e) = ve, L(a_i) = vi, L(a_c) = vc, L(bad) = nb;
        END_LANES
        double r0[NX], x0[NX], e[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) r0[i] = w.RG[NU + i], x0[i] = w.Z[NU + i];
        proj0(r0);
        eq0_violation(x0, e);
        rs = WARP_MAX(a_s), re = WARP_MAX(a_e), ri = WARP_MAX(a_i), rc = WARP_MAX(a_c);
        bool nan = WARP_ANY(bad);
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            nan |= (r0[i] != r0[i]);
            rs = fmax(rs, fabs(r0[i]));
            ri = fmax(ri, fabs(e[i]));
        }
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
            w.LAMQ[k * NC + i] = ll, w.LAMQ[k * NC + NZ + i] = lu;
            w.TQ[k * NC + i] = tl, w.TQ[k * NC + NZ + i] = tu;
        }
        for (int idx = lane; idx < N * NX; idx += 32) w.PIQ[idx] = 0.0;
        END_LANES
        double x0[NX], e[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) x0[i] = w.Z[NU + i] + w.DZ[NU + i];
        eq0_violation(x0, e);
        UNIFORM_SYNC();  // the loads above precede the stores below
        FOR_LANES
        if (lane < NX) {
            double ev = 0.0;
#pragma unroll
            for (int i = 0; i < NX; ++i)
                if (i == lane) ev = e[i];
            w.DZ[NU + lane] -= ev;
        }
        END_LANES
    }

    // ---------------------------------------------------------------- QP: residuals
    // RG (stationarity), RB (dynamics), RD (bounds), RMB (lam*t); returns mu, norms by reference.
    VB_DEV double qp_residuals(double &ng, double &nb_, double &nd, double &nm, bool &nan) {
        const int N = s.N;
        LV(double, a_g);
        LV(double, a_b);
        LV(double, a_d);
        LV(double, a_m);
        LV(double, a_mu);
        LV(int, bad);
        FOR_LANES
        double vg = 0, vb = 0, vd = 0, vm = 0, mu = 0;
        int nb = 0;
        for (int idx = lane; idx < (N + 1) * NZ; idx += 32) {
            int k = idx / NZ, i = idx - k * NZ;
            double v = w.DZ[idx];
            double r = cost_h(k, i) * v + cost_g(k, i, w.Z[idx]);
            if (k < N) {
                const double *col = w.BAT + (size_t)idx * NX, *pi = w.PIQ + k * NX;
#pragma unroll
                for (int m = 0; m < NX; ++m) r += col[m] * pi[m];
            }
            if (k > 0 && i >= NU) r -= w.PIQ[(k - 1) * NX + i - NU];
            int c = k * NC + i;
            if (active(k, i)) {
                int sc = sclass(k);
                double z = w.Z[idx];
                double ll = w.LAMQ[c], lu = w.LAMQ[c + NZ], tl = w.TQ[c], tu = w.TQ[c + NZ];
                r += lu - ll;
                double dl = (s.lb[sc][i] - z) - v + tl, du = v - (s.ub[sc][i] - z) + tu;
                double ml = ll * tl, mu_ = lu * tu;
                w.RD[c] = dl, w.RD[c + NZ] = du, w.RMB[c] = ml, w.RMB[c + NZ] = mu_;
                nb |= (dl != dl) | (du != du) | (ml != ml) | (mu_ != mu_);
                vd = fmax(vd, fmax(fabs(dl), fabs(du)));
                vm = fmax(vm, fmax(fabs(ml), fabs(mu_)));
                mu += ml + mu_;
            } else {
                w.RD[c] = 0.0, w.RD[c + NZ] = 0.0, w.RMB[c] = 0.0, w.RMB[c + NZ] = 0.0;
                if (k == N) r = 0.0;
            }
            w.RG[idx] = r;
            if (!(k == 0 && i >= NU)) {
                nb |= (r != r);
                vg = fmax(vg, fabs(r));
            }
        }
        for (int idx = lane; idx < N * NX; idx += 32) {
            int k = idx / NX, i = idx - k * NX;
            double a = w.BD[idx] - w.DZ[(k + 1) * NZ + NU + i];
            const double *dz = w.DZ + k * NZ, *bat = w.BAT + (size_t)k * NZ * NX + i;
#pragma unroll
            for (int j = 0; j < NZ; ++j) a += bat[j * NX] * dz[j];
            w.RB[idx] = a;
            nb |= (a != a);
            vb = fmax(vb, fabs(a));
        }
        L(a_g) = vg, L(a_b) = vb, L(a_d) = vd, L(a_m) = vm, L(a_mu) = mu, L(bad) = nb;
        END_LANES
        double r0[NX], x0[NX], e[NX], eN[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            r0[i] = w.RG[NU + i];
            x0[i] = w.Z[NU + i] + w.DZ[NU + i];
            eN[i] = ((s.fixedN >> i) & 1) ? w.Z[N * NZ + NU + i] + w.DZ[N * NZ + NU + i] - s.cN[i] : 0.0;
        }
        proj0(r0);
        eq0_violation(x0, e);
        ng = WARP_MAX(a_g), nb_ = WARP_MAX(a_b), nd = WARP_MAX(a_d), nm = WARP_MAX(a_m);
        double mu = WARP_SUM(a_mu);
        nan = WARP_ANY(bad);
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            nan |= (r0[i] != r0[i]);
            ng = fmax(ng, fabs(r0[i]));
            nb_ = fmax(nb_, fmax(fabs(e[i]), fabs(eN[i])));
        }
        UNIFORM_SYNC();
        FOR_LANES
        if (lane < NX) {
            double rv = 0, ev = 0, env = 0;
#pragma unroll
            for (int i = 0; i < NX; ++i)
                if (i == lane) rv = r0[i], ev = e[i], env = eN[i];
            w.RG[NU + lane] = rv;
            s.e0[lane] = ev, s.eN[lane] = env;
        }
        END_LANES
        return s.nact ? mu / (2.0 * s.nact) : 0.0;
    }

    // effective Hessian diagonal / gradient of component (k, i) for the Newton system
    VB_DEV void eff(int k, int i, const double *RMs, double &hh, double &rr) const {
        hh = cost_h(k, i) + o.qp_reg_prim;
        rr = w.RG[k * NZ + i];
        if (active(k, i)) {
            int c = k * NC + i;
            double ll = w.LAMQ[c], lu = w.LAMQ[c + NZ], tl = w.TQ[c], tu = w.TQ[c + NZ];
            hh += ll / tl + lu / tu;
            rr += (RMs[c] - ll * w.RD[c]) / tl - (RMs[c + NZ] - lu * w.RD[c + NZ]) / tu;
        }
    }

    // ---------------------------------------------------------------- Riccati: backward sweep
    // factor = true: factorise and solve; false: re-use the stored factors with a new rhs.
    // Ends with the stage-0 step in s.dz[NU..].  Returns false on a singular terminal block / NaN.
    VB_DEV bool backward(bool factor, const double *RMs) {
        const int N = s.N;
        FOR_LANES
        if (lane < NX) {
            double hh, rr;
            eff(N, NU + lane, RMs, hh, rr);
            bool fx = (s.fixedN >> lane) & 1;
            s.hhN[lane] = fx ? 0.0 : hh;
            s.rN[lane] = fx ? 0.0 : rr;
            s.pvec[lane] = fx ? 0.0 : rr;
#pragma unroll
            for (int j = 0; j < NX; ++j) s.P[lane][j] = (j == lane && !fx) ? hh : 0.0;
        }
        END_LANES
        bool ok = true;
        for (int k = N - 1; k >= 0; --k) {
            const bool last = (k == N - 1) && s.termfix;
            double *fac = w.FAC + (size_t)k * FS;
            // A: stage data to shared memory
            FOR_LANES
            for (int idx = lane; idx < NZ * NX; idx += 32)
                (&s.BAT[0][0])[idx] = w.BAT[(size_t)k * NZ * NX + idx];
            if (lane < NX) s.beta[lane] = w.RB[k * NX + lane];
            if (lane < NZ) {
                double hh, rr;
                eff(k, lane, RMs, hh, rr);
                s.hh[lane] = hh, s.rr[lane] = rr;
            }
            if (!factor) {
                if (k < N - 1) {
                    const double *Pn = w.FAC + (size_t)(k + 1) * FS + OFF_P;
                    for (int idx = lane; idx < NX * NX; idx += 32) (&s.P[0][0])[idx] = Pn[idx];
                }
                if (last)
                    for (int idx = lane; idx < NZ * NZ; idx += 32) (&s.M[0][0])[idx] = w.MF[idx];
            }
            END_LANES
            // B: P+ [B A]  and  t2 = P+ beta + p+
            FOR_LANES
            if (factor) {
                for (int idx = lane; idx < NZ * NX; idx += 32) {
                    int j = idx / NX, i = idx - j * NX;
                    double a = 0.0;
#pragma unroll
                    for (int m = 0; m < NX; ++m) a += s.P[i][m] * s.BAT[j][m];
                    s.PBAT[j][i] = a;
                }
            }
            if (lane < NX) {
                double a = s.pvec[lane];
#pragma unroll
                for (int m = 0; m < NX; ++m) a += s.P[lane][m] * s.beta[m];
                s.t2[lane] = a;
            }
            END_LANES
            // C: M = H + [B A]' P+ [B A]  and  m = r + [B A]' t2
            FOR_LANES
            if (factor) {
                for (int idx = lane; idx < TRI; idx += 32) {
                    int a_ = 0;
                    while ((a_ + 1) * (a_ + 2) / 2 <= idx) ++a_;
                    int b_ = idx - a_ * (a_ + 1) / 2;
                    double a = (a_ == b_) ? s.hh[a_] : 0.0;
#pragma unroll
                    for (int m = 0; m < NX; ++m) a += s.BAT[a_][m] * s.PBAT[b_][m];
                    s.M[a_][b_] = a, s.M[b_][a_] = a;
                }
            }
            if (lane < NZ) {
                double a = s.rr[lane];
#pragma unroll
                for (int i = 0; i < NX; ++i) a += s.BAT[lane][i] * s.t2[i];
                s.m[lane] = a;
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
                        di[j] = d > 0.0 ? 1.0 / sqrt(d) : 0.0;
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
                        di[i] = fac[i * NU + i];
#pragma unroll
                        for (int c = 0; c < i; ++c) Lu[i][c] = fac[i * NU + c];
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
                    for (int idx = lane; idx < NX * NX; idx += 32) {
                        int i = idx / NX, j = idx - i * NX;
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
                        s.P[i][j] = a;
                        fac[OFF_P + idx] = a;
                        if (j == 0) {
#pragma unroll
                            for (int c = 0; c < NU; ++c) fac[OFF_LXU + i * NU + c] = li[c];
                        }
                    }
                    if (lane == 0) {
#pragma unroll
                        for (int i = 0; i < NU; ++i)
#pragma unroll
                            for (int c = 0; c < NU; ++c)
                                fac[i * NU + c] = (c == i) ? di[i] : (c < i ? Lu[i][c] : 0.0);
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
                        for (int c = 0; c < NU; ++c) li[c] = fac[OFF_LXU + lane * NU + c];
                    }
                    double p = s.m[NU + lane];
#pragma unroll
                    for (int c = 0; c < NU; ++c) p -= li[c] * y[c];
                    s.pvec[lane] = p;
                    w.PV[k * NX + lane] = p;
                }
                if (lane < NU) {
                    double yv = 0.0;
#pragma unroll
                    for (int c = 0; c < NU; ++c)
                        if (c == lane) yv = y[c];
                    w.YV[k * NU + lane] = yv;
                }
                END_LANES
            } else {
                // terminal equalities: G du + Gx dx + beta_v = -eN on the velocity rows determines
                // du = K dx + k0; the value function is the restriction of M to that manifold.
                double Gi[NU][NU], K[NU][NX], k0[NU], tmp[NZ];
                if (factor) {
                    double G[NU][NU];
#pragma unroll
                    for (int a = 0; a < NU; ++a)
#pragma unroll
                        for (int b = 0; b < NU; ++b) G[a][b] = s.BAT[b][NQ + a];
                    ok = inverse_small(G, Gi) && ok;
#pragma unroll
                    for (int a = 0; a < NU; ++a)
#pragma unroll
                        for (int j = 0; j < NX; ++j) {
                            double v = 0.0;
#pragma unroll
                            for (int b = 0; b < NU; ++b) v -= Gi[a][b] * s.BAT[NU + j][NQ + b];
                            K[a][j] = v;
                        }
                } else {
#pragma unroll
                    for (int a = 0; a < NU; ++a) {
#pragma unroll
                        for (int b = 0; b < NU; ++b) Gi[a][b] = fac[a * NU + b];
#pragma unroll
                        for (int j = 0; j < NX; ++j) K[a][j] = fac[OFF_LXU + j * NU + a];
                    }
                }
#pragma unroll
                for (int a = 0; a < NU; ++a) {
                    double v = 0.0;
#pragma unroll
                    for (int b = 0; b < NU; ++b) v -= Gi[a][b] * (s.eN[NQ + b] + s.beta[NQ + b]);
                    k0[a] = v;
                }
#pragma unroll
                for (int i = 0; i < NZ; ++i) {
                    double v = s.m[i];
#pragma unroll
                    for (int a = 0; a < NU; ++a) v += s.M[i][a] * k0[a];
                    tmp[i] = v;
                }
                FOR_LANES
                if (factor) {
                    for (int idx = lane; idx < NX * NX; idx += 32) {
                        int i = idx / NX, j = idx - i * NX;
                        double ki[NU], kj[NU];
#pragma unroll
                        for (int a = 0; a < NU; ++a) {
                            double x1 = 0, x2 = 0;
#pragma unroll
                            for (int jj = 0; jj < NX; ++jj) {
                                if (jj == i) x1 = K[a][jj];
                                if (jj == j) x2 = K[a][jj];
                            }
                            ki[a] = x1, kj[a] = x2;
                        }
                        double a = s.M[NU + i][NU + j];
#pragma unroll
                        for (int c = 0; c < NU; ++c) {
                            a += ki[c] * s.M[c][NU + j] + s.M[NU + i][c] * kj[c];
#pragma unroll
                            for (int c2 = 0; c2 < NU; ++c2) a += ki[c] * s.M[c][c2] * kj[c2];
                        }
                        fac[OFF_P + idx] = a;
                        if (j == 0) {
#pragma unroll
                            for (int c = 0; c < NU; ++c) fac[OFF_LXU + i * NU + c] = ki[c];
                        }
                        s.P[i][j] = a;
                    }
                    for (int idx = lane; idx < NZ * NZ; idx += 32) w.MF[idx] = (&s.M[0][0])[idx];
                    if (lane == 0) {
#pragma unroll
                        for (int a = 0; a < NU; ++a)
#pragma unroll
                            for (int b = 0; b < NU; ++b) fac[a * NU + b] = Gi[a][b];
                    }
                }
                if (lane < NX) {
                    double p = 0.0;
#pragma unroll
                    for (int jj = 0; jj < NX; ++jj)
                        if (jj == lane) {
                            p = tmp[NU + jj];
#pragma unroll
                            for (int a = 0; a < NU; ++a) p += K[a][jj] * tmp[a];
                        }
                    s.pvec[lane] = p;
                    w.PV[k * NX + lane] = p;
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
        }
        // stage 0:  dx0 = -e0 + Z0 dy,  (Z0'P0 Z0) dy = -Z0'(p0 - P0 e0)
        {
            const double *P0 = w.FAC + OFF_P;
            double Lz[NX][NX], dzi[NX], Pe[NX], rhs[NX], y[NX], dy[NX];
            if (factor) {
                double T[NX][NX];  // P0 Z0
#pragma unroll
                for (int i = 0; i < NX; ++i)
#pragma unroll
                    for (int c = 0; c < NX; ++c) {
                        double a = 0.0;
#pragma unroll
                        for (int j = 0; j < NX; ++j) a += P0[i * NX + j] * s.Z0[j][c];
                        T[i][c] = a;
                    }
#pragma unroll
                for (int a_ = 0; a_ < NX; ++a_)
#pragma unroll
                    for (int b_ = 0; b_ <= a_; ++b_) {
                        double a = 0.0;
#pragma unroll
                        for (int i = 0; i < NX; ++i) a += s.Z0[i][a_] * T[i][b_];
                        Lz[a_][b_] = a;
                    }
#pragma unroll
                for (int j = 0; j < NX; ++j) {
                    double d = Lz[j][j];
#pragma unroll
                    for (int c = 0; c < j; ++c) d -= Lz[j][c] * Lz[j][c];
                    dzi[j] = d > 0.0 ? 1.0 / sqrt(d) : 0.0;
#pragma unroll
                    for (int i = j + 1; i < NX; ++i) {
                        double a = Lz[i][j];
#pragma unroll
                        for (int c = 0; c < j; ++c) a -= Lz[i][c] * Lz[j][c];
                        Lz[i][j] = a * dzi[j];
                    }
                    Lz[j][j] = dzi[j];
                }
            } else {
#pragma unroll
                for (int i = 0; i < NX; ++i) {
#pragma unroll
                    for (int c = 0; c <= i; ++c) Lz[i][c] = s.Lz[i][c];
                    dzi[i] = Lz[i][i];
                }
            }
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                double a = 0.0;
#pragma unroll
                for (int j = 0; j < NX; ++j) a += P0[i * NX + j] * s.e0[j];
                Pe[i] = a;
            }
#pragma unroll
            for (int c = 0; c < NX; ++c) {
                double a = 0.0;
#pragma unroll
                for (int i = 0; i < NX; ++i) a -= s.Z0[i][c] * (s.pvec[i] - Pe[i]);
                rhs[c] = a;
            }
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                double a = rhs[i];
#pragma unroll
                for (int j = 0; j < i; ++j) a -= Lz[i][j] * y[j];
                y[i] = a * dzi[i];
            }
#pragma unroll
            for (int i = NX - 1; i >= 0; --i) {
                double a = y[i];
#pragma unroll
                for (int j = i + 1; j < NX; ++j) a -= Lz[j][i] * dy[j];
                dy[i] = a * dzi[i];
            }
            FOR_LANES
            if (lane < NX) {
                double a = 0.0;
#pragma unroll
                for (int i = 0; i < NX; ++i)
                    if (i == lane) {
                        a = -s.e0[i];
#pragma unroll
                        for (int c = 0; c < NX; ++c) a += s.Z0[i][c] * dy[c];
                    }
                s.dz[NU + lane] = a;
                if (factor) {
#pragma unroll
                    for (int i = 0; i < NX; ++i)
                        if (i == lane) {
#pragma unroll
                            for (int c = 0; c <= i; ++c) s.Lz[i][c] = Lz[i][c];
                        }
                }
            }
            END_LANES
        }
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

    // step of one bound constraint and its contribution to the maximum step length
    VB_DEV void con_step(int k, int c, double dvv, const double *RMs, double &amin) {
        int sgn = c >= NZ, i = c - sgn * NZ;
        if (!active(k, i)) return;
        int cc = k * NC + c;
        double lam = w.LAMQ[cc], t = w.TQ[cc];
        double dtt = (sgn ? -dvv : dvv) - w.RD[cc];
        double dl = -(RMs[cc] + lam * dtt) / t;
        w.DT[cc] = dtt, w.DLAM[cc] = dl;
        if (dtt < 0.0) amin = fmin(amin, -t / dtt);
        if (dl < 0.0) amin = fmin(amin, -lam / dl);
    }

    // ---------------------------------------------------------------- Riccati: forward sweep
    // DV, DPI, DT, DLAM from the factors; returns the maximum step to the boundary.
    VB_DEV double forward(const double *RMs) {
        const int N = s.N;
        LV(double, amin);
        FOR_LANES
        L(amin) = 1.0;
        END_LANES
        for (int k = 0; k < N; ++k) {
            const bool last = (k == N - 1) && s.termfix;
            const double *fac = w.FAC + (size_t)k * FS;
            FOR_LANES
            for (int idx = lane; idx < NZ * NX; idx += 32)
                (&s.BAT[0][0])[idx] = w.BAT[(size_t)k * NZ * NX + idx];
            if (lane < NX) s.beta[lane] = w.RB[k * NX + lane];
            if (k + 1 < N) {
                const double *Pn = w.FAC + (size_t)(k + 1) * FS + OFF_P;
                for (int idx = lane; idx < NX * NX; idx += 32) (&s.P[0][0])[idx] = Pn[idx];
                if (lane < NX) s.pvec[lane] = w.PV[(k + 1) * NX + lane];
            } else if (lane < NX) {
#pragma unroll
                for (int j = 0; j < NX; ++j) s.P[lane][j] = (j == lane) ? s.hhN[lane] : 0.0;
                s.pvec[lane] = s.rN[lane];
            }
            END_LANES
            double du[NU];
            if (!last) {
                double t[NU];
#pragma unroll
                for (int c = 0; c < NU; ++c) {
                    double a = w.YV[k * NU + c];
#pragma unroll
                    for (int j = 0; j < NX; ++j) a += fac[OFF_LXU + j * NU + c] * s.dz[NU + j];
                    t[c] = a;
                }
#pragma unroll
                for (int c = NU - 1; c >= 0; --c) {
                    double a = -t[c];
#pragma unroll
                    for (int c2 = c + 1; c2 < NU; ++c2) a -= fac[c2 * NU + c] * du[c2];
                    du[c] = a * fac[c * NU + c];
                }
            } else {
#pragma unroll
                for (int a_ = 0; a_ < NU; ++a_) {
                    double a = w.MF[NZ * NZ + NZ + a_];
#pragma unroll
                    for (int j = 0; j < NX; ++j) a += fac[OFF_LXU + j * NU + a_] * s.dz[NU + j];
                    du[a_] = a;
                }
            }
            FOR_LANES
            if (lane < NU) {
                double v = 0.0;
#pragma unroll
                for (int c = 0; c < NU; ++c)
                    if (c == lane) v = du[c];
                s.dz[lane] = v;
                w.DV[k * NZ + lane] = v;
            } else if (lane < NZ) {
                w.DV[k * NZ + lane] = s.dz[lane];
            }
            END_LANES
            FOR_LANES
            if (lane < NX) {
                double a = s.beta[lane];
#pragma unroll
                for (int j = 0; j < NZ; ++j) a += s.BAT[j][lane] * s.dz[j];
                s.dxn[lane] = a;
            } else if (lane < NX + NC) {
                int c = lane - NX;
                con_step(k, c, s.dz[c >= NZ ? c - NZ : c], RMs, L(amin));
            }
            END_LANES
            double nuv[NU];
            if (last) {
                // multiplier of the eliminated velocity rows from the u-stationarity of the last stage
                double tu[NU];
#pragma unroll
                for (int a_ = 0; a_ < NU; ++a_) {
                    double a = w.MF[NZ * NZ + a_];
#pragma unroll
                    for (int j = 0; j < NZ; ++j) a += w.MF[a_ * NZ + j] * s.dz[j];
                    tu[a_] = a;
                }
#pragma unroll
                for (int b = 0; b < NU; ++b) {
                    double a = 0.0;
#pragma unroll
                    for (int a_ = 0; a_ < NU; ++a_) a -= fac[a_ * NU + b] * tu[a_];
                    nuv[b] = a;
                }
            }
            UNIFORM_SYNC();
            FOR_LANES
            if (lane < NX) {
                double a = s.pvec[lane];
#pragma unroll
                for (int m = 0; m < NX; ++m) a += s.P[lane][m] * s.dxn[m];
                if (last && lane >= NQ) {
#pragma unroll
                    for (int b = 0; b < NU; ++b)
                        if (b == lane - NQ) a = nuv[b];
                }
                w.DPI[k * NX + lane] = a;
                s.dz[NU + lane] = s.dxn[lane];
            }
            END_LANES
        }
        FOR_LANES
        if (lane < NZ) w.DV[N * NZ + lane] = lane < NU ? 0.0 : s.dz[lane];
        for (int c = lane; c < NC; c += 32) {
            int i = c >= NZ ? c - NZ : c;
            con_step(N, c, i < NU ? 0.0 : s.dz[i], RMs, L(amin));
        }
        END_LANES
        return WARP_MIN(amin);
    }

    VB_DEV double mu_at(double alpha) {
        const int N = s.N;
        LV(double, acc);
        FOR_LANES
        double a = 0.0;
        for (int idx = lane; idx < (N + 1) * NC; idx += 32) {
            int k = idx / NC, c = idx - k * NC, i = c >= NZ ? c - NZ : c;
            if (active(k, i)) a += (w.LAMQ[idx] + alpha * w.DLAM[idx]) * (w.TQ[idx] + alpha * w.DT[idx]);
        }
        L(acc) = a;
        END_LANES
        double tot = WARP_SUM(acc);
        return s.nact ? tot / (2.0 * s.nact) : 0.0;
    }

    // ---------------------------------------------------------------- IPM
    // returns 0 success, 1 max iter, 2 min step, 3 NaN
    VB_DEV int ipm_solve(int &iters) {
        const int N = s.N;
        qp_init();
        double rg, rb, rd, rm, alpha = 1.0;
        bool nan = false, ok = true;
        double mu = qp_residuals(rg, rb, rd, rm, nan);
        int kk = 0;
        for (; kk < o.qp_iter_max && alpha > o.qp_alpha_min && !nan &&
               (rg > o.qp_tol_stat || rb > o.qp_tol_eq || rd > o.qp_tol_ineq || rm > o.qp_tol_comp);
             ++kk) {
            // predictor: res_m = lam * t
            ok = backward(true, w.RMB);
            if (!ok) break;
            double a_aff = forward(w.RMB);
            double m_aff = mu_at(a_aff);
            double sigma = m_aff / mu;
            sigma = sigma * sigma * sigma;
            double sm = sigma * mu;
            if (sm < o.qp_tau_min) sm = o.qp_tau_min;
            // centering + corrector: res_m = lam*t + dt_aff*dlam_aff - sigma*mu
            FOR_LANES
            for (int idx = lane; idx < (N + 1) * NC; idx += 32) {
                int k = idx / NC, c = idx - k * NC, i = c >= NZ ? c - NZ : c;
                w.RM[idx] = active(k, i) ? w.RMB[idx] + w.DT[idx] * w.DLAM[idx] - sm : 0.0;
            }
            END_LANES
            backward(false, w.RM);
            alpha = forward(w.RM);
            // conditional predictor-corrector
            double m_cor = mu_at(alpha);
            if (m_cor > 2.0 * m_aff) {
                FOR_LANES
                for (int idx = lane; idx < (N + 1) * NC; idx += 32) {
                    int k = idx / NC, c = idx - k * NC, i = c >= NZ ? c - NZ : c;
                    w.RM[idx] = active(k, i) ? w.RMB[idx] - sm : 0.0;
                }
                END_LANES
                backward(false, w.RM);
                alpha = forward(w.RM);
            }
            double as = alpha;
            if (as < 1.0) as = as * ((1.0 - as) * 0.99 + as * 0.9999);
            FOR_LANES
            for (int idx = lane; idx < (N + 1) * NZ; idx += 32) w.DZ[idx] += as * w.DV[idx];
            for (int idx = lane; idx < N * NX; idx += 32) w.PIQ[idx] += as * w.DPI[idx];
            for (int idx = lane; idx < (N + 1) * NC; idx += 32) {
                int k = idx / NC, c = idx - k * NC, i = c >= NZ ? c - NZ : c;
                if (active(k, i)) {
                    w.LAMQ[idx] = fmax(w.LAMQ[idx] + as * w.DLAM[idx], o.qp_lam_min);
                    w.TQ[idx] = fmax(w.TQ[idx] + as * w.DT[idx], o.qp_t_min);
                }
            }
            END_LANES
            mu = qp_residuals(rg, rb, rd, rm, nan);
        }
        iters = kk;
        // multipliers of the eliminated equalities from stationarity
        {
            double r[NX], rp[NX], nuN[NX];
#pragma unroll
            for (int i = 0; i < NX; ++i) {
                double a = cost_h(0, NU + i) * w.DZ[NU + i] + cost_g(0, NU + i, w.Z[NU + i]);
                const double *col = w.BAT + (size_t)(NU + i) * NX;
#pragma unroll
                for (int m = 0; m < NX; ++m) a += col[m] * w.PIQ[m];
                if (active(0, NU + i)) a += w.LAMQ[NZ + NU + i] - w.LAMQ[NU + i];
                r[i] = rp[i] = a;
                nuN[i] = ((s.fixedN >> i) & 1)
                             ? w.PIQ[(N - 1) * NX + i] - cost_g(N, NU + i, w.Z[N * NZ + NU + i]) -
                                   cost_h(N, NU + i) * w.DZ[N * NZ + NU + i]
                             : 0.0;
            }
            proj0(rp);
            FOR_LANES
            if (lane < NX) {
#pragma unroll
                for (int i = 0; i < NX; ++i)
                    if (i == lane) s.nu0q[lane] = r[i] - rp[i], s.nuNq[lane] = nuN[i];
            }
            END_LANES
        }
        if (!ok || nan || mu != mu) return 3;
        if (rg > o.qp_tol_stat || rb > o.qp_tol_eq || rd > o.qp_tol_ineq || rm > o.qp_tol_comp)
            return kk >= o.qp_iter_max ? 1 : 2;
        return 0;
    }

    // ---------------------------------------------------------------- merit function
    VB_DEV double total_cost(const double *Zs) {
        const int N = s.N;
        if (FAM == VBOC_FAMILY_VBOC) {
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
#pragma unroll
            for (int i = 0; i < NQ; ++i) q += Zs[k * NZ + NU + NQ + i] * Zs[k * NZ + NU + NQ + i];
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
                if (fl > 0) a += w.WB[k * NC + i] * fl;
                if (fu > 0) a += w.WB[k * NC + NZ + i] * fu;
            }
        }
        L(acc) = a;
        END_LANES
        double m = WARP_SUM(acc) + total_cost(Zs);
        double x0[NX], e[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) x0[i] = Zs[NU + i];
        eq0_violation(x0, e);
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            m += s.w0[i] * fabs(e[i]);
            if ((s.fixedN >> i) & 1) m += s.wN[i] * fabs(Zs[N * NZ + NU + i] - s.cN[i]);
        }
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
        END_LANES
        double m0 = merit(w.Z);
        double alpha = 1.0;
        for (;;) {
            FOR_LANES
            for (int idx = lane; idx < (N + 1) * NZ; idx += 32) w.ZT[idx] = w.Z[idx] + alpha * w.DZ[idx];
            END_LANES
            double m1 = merit(w.ZT);
            ++evals;
            if (m1 < m0) break;
            if (alpha * o.alpha_reduction < o.alpha_min) break;  // the smallest step is taken anyway
            alpha *= o.alpha_reduction;
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
            END_LANES
        }
        st.cost = total_cost(w.Z);
        store_solution(pb, st);
    }
};

}  // namespace vboc
