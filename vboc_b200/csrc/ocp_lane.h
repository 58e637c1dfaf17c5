// ocp_lane.h -- one LANE solves one OCP (32 OCPs per warp): the same SQP / Mehrotra IPM / Riccati algorithm
// as ocp_warp.h, written as plain per-thread code.
//
// Why a second mapping: with blocks of at most 9 x 9 the warp-per-OCP kernel spends ~130 k warp
// instructions per IPM iteration (every FMA fetches its operands from shared memory, most lanes idle in the
// serial sweeps; profiles/r1_solve_kernel_summary.md) where the arithmetic needs ~6 k.  Here every lane
// keeps the stage blocks in registers, the 32 lanes of a warp advance 32 OCPs in lockstep (SIMT handles
// the differing IPM iteration counts: finished lanes idle until the warp's slowest QP is done), and the
// workspace is interleaved [stage][field][lane] so that every load / store of the warp is one coalesced
// 256-byte line.  Finished lanes pull the next problem at SQP-iteration boundaries (solve_lockstep in
// vboc_cuda.cu), so SQP iteration counts do not have to agree within a warp.
//
// W = 32 on the GPU; W = 1 compiles the same code for the host (tools/emu) for validation against the oracle.
#pragma once
#include <stddef.h>

#include "../../include/vboc_b200.h"
#include "dynamics.h"
#include "ocp_warp.h"  // Prob, Dim

namespace vboc {

// DTS = 1: the dt state of the VBOC models is kept (free dt, VBOC/pendulum_class_vboc.py); nx = 2n + 1 and the
// dynamics are the dt-scaled ones over a unit step.  Terminal equalities that the last control cannot absorb
// (there: theta_N and dtheta_N fixed with one control) are handled by BORDERING: the Riccati system is solved
// for the rhs and for one unit terminal gradient per fixed component, and the terminal multipliers follow
// from a small dense system.
template <int NQ, int DTS = 0>
struct LaneLayout {
    static constexpr int NX = 2 * NQ + DTS, NU = NQ, NZ = NX + NU, NC = 2 * NZ;
    static constexpr int NB = DTS ? NX : 0;  // room for the border basis solves
    // per stage (element offsets)
    static constexpr size_t Z = 0, PI = Z + NZ, LAM = PI + NX, BD = LAM + NC;
    static constexpr size_t BA = BD + NX, RB = BA + NX * NZ, HH = RB + NX, RR = HH + NZ, Q1 = RR + NZ, Q2 = Q1 + NZ;
    static constexpr size_t MB = Q2 + NZ, LUU = MB + NZ, LXU = LUU + NU * NU, YV = LXU + NX * NU, P = YV + NU;
    static constexpr size_t PV = P + NX * NX, DZ = PV + NX, PIQ = DZ + NZ, LAMQ = PIQ + NX, TQ = LAMQ + NC;
    static constexpr size_t DV = TQ + NC, DPI = DV + NZ, PROD = DPI + NX, DLAM = PROD + NC, DT = DLAM + NC;
    static constexpr size_t WDYN = DT + NC, WB = WDYN + NX, ZT = WB + NC, DVB = ZT + NZ, DPIB = DVB + NB * NZ;
    static constexpr size_t SREC = DPIB + NB * NX;
    static constexpr size_t SMAX = 129;
    // per OCP
    static constexpr size_t O0 = SMAX * SREC;
    static constexpr size_t LB = O0, UB = LB + 3 * NZ, Z0 = UB + 3 * NZ, C0 = Z0 + NX * NX, CN = C0 + NX;
    static constexpr size_t MF = CN + NX, ML = MF + NZ * NZ, K0 = ML + NZ, LZ = K0 + NU, DZI = LZ + NX * NX;
    static constexpr size_t PE = DZI + NX, E0 = PE + NX, EN = E0 + NX, HHN = EN + NX, RN = HHN + NX, NUV = RN + NX;
    static constexpr size_t W0 = NUV + NX, WN = W0 + NX, NU0Q = WN + NX, NUNQ = NU0Q + NX, TOTAL = NUNQ + NX;
};

// persistent per-lane state of the SQP loop (registers)
struct LaneState {
    int have;  // a problem is loaded
    int it;    // SQP iteration counter
    vboc_stats st;
};

template <int NQ, int FAM, int W, int DTS = 0>
struct LaneSolver {
    using L = LaneLayout<NQ, DTS>;
    static constexpr int NX = L::NX, NU = L::NU, NZ = L::NZ, NC = L::NC;

    double *base;
    int lane;
    const vboc_opts &o;
    // per-OCP scalars
    int N, fixed0, fixedN, termfix, nact;
    int nb, bidx[NX];  // border mode: number and indices of the fixed terminal components
    double h, wtdt, wt, wcost[NQ];

    VB_HD LaneSolver(double *base_, int lane_, const vboc_opts &o_) : base(base_), lane(lane_), o(o_), N(0) {}

    VB_HD double &g(size_t off) const { return base[off * W + lane]; }
    VB_HD double &s(int k, size_t f) const { return base[((size_t)k * L::SREC + f) * W + lane]; }

    VB_HD int sclass(int k) const { return k == 0 ? 0 : (k == N ? 2 : 1); }
    VB_HD bool active(int k, int i) const {
        if (k == N) {
            if (i < NU) return false;
            if ((fixedN >> (i - NU)) & 1) return false;
        }
        if (k == 0 && i >= NU && ((fixed0 >> (i - NU)) & 1)) return false;
        return true;
    }
    VB_HD double lb(int k, int i) const { return g(L::LB + sclass(k) * NZ + i); }
    VB_HD double ub(int k, int i) const { return g(L::UB + sclass(k) * NZ + i); }
    VB_HD double cost_g(int k, int i, double zval) const {
        if (FAM == VBOC_FAMILY_VBOC) {
            if (k == 0 && i >= NU + NQ && i < NU + 2 * NQ) return wcost[i - NU - NQ];
            if (DTS && i == NU + 2 * NQ && k < N) return wt;  // wt * dt at stages 0..N-1
            return 0.0;
        }
        return (i >= NU + NQ) ? 2.0 * (k < N ? h : 1.0) * zval : 0.0;
    }
    VB_HD double cost_h(int k, int i) const {
        double hd = o.levenberg_marquardt;
        if (FAM == VBOC_FAMILY_AL && i >= NU + NQ) hd += 2.0 * (k < N ? h : 1.0);
        return hd;
    }
    // e = (I - Z0 Z0')(x - c0);  v <- Z0 Z0' v
    VB_HD void eq0_violation(const double *x, double *e) const {
        double y[NX];
        for (int c = 0; c < NX; ++c) {
            double a = 0.0;
            for (int i = 0; i < NX; ++i) a += g(L::Z0 + i * NX + c) * (x[i] - g(L::C0 + i));
            y[c] = a;
        }
        for (int i = 0; i < NX; ++i) {
            double a = x[i] - g(L::C0 + i);
            for (int c = 0; c < NX; ++c) a -= g(L::Z0 + i * NX + c) * y[c];
            e[i] = a;
        }
    }
    VB_HD void proj0(double *v) const {
        double y[NX], out[NX];
        for (int c = 0; c < NX; ++c) {
            double a = 0.0;
            for (int i = 0; i < NX; ++i) a += g(L::Z0 + i * NX + c) * v[i];
            y[c] = a;
        }
        for (int i = 0; i < NX; ++i) {
            double a = 0.0;
            for (int c = 0; c < NX; ++c) a += g(L::Z0 + i * NX + c) * y[c];
            out[i] = a;
        }
        for (int i = 0; i < NX; ++i) v[i] = out[i];
    }

    // ---------------------------------------------------------------- problem load / store
    VB_HD void load_problem(const Prob &pb) {
        N = pb.N, h = pb.h;
        wt = pb.wt;
        wtdt = (FAM == VBOC_FAMILY_VBOC && !DTS) ? pb.wt * pb.h * N : 0.0;
        int f0 = 0, fN = 0, nf0 = 0, nfN = 0;
        for (int i = 0; i < NX; ++i) {
            bool a = pb.lbx0[i] == pb.ubx0[i], b = pb.lbxN[i] == pb.ubxN[i];
            f0 |= (int)a << i, fN |= (int)b << i, nf0 += a, nfN += b;
            g(L::C0 + i) = a ? pb.lbx0[i] : 0.0;
            g(L::CN + i) = b ? pb.lbxN[i] : 0.0;
            for (int c = 0; c < NX; ++c) g(L::Z0 + i * NX + c) = 0.0;
            g(L::W0 + i) = 0.0, g(L::WN + i) = 0.0;
        }
        int ny = 0;
        for (int i = 0; i < NX; ++i) {
            if ((f0 >> i) & 1) continue;
            if (pb.dir && i >= NQ && i < 2 * NQ) continue;
            g(L::Z0 + i * NX + ny++) = 1.0;
        }
        if (pb.dir) {
            for (int i = 0; i < NQ; ++i) g(L::Z0 + (NQ + i) * NX + ny) = pb.dir[i];
            ++ny;
        }
        fixed0 = f0, fixedN = fN;
        // the last control absorbs the terminal equalities iff they are exactly the n velocities
        termfix = fN == (((1 << NQ) - 1) << NQ);
        nb = 0;
        if (!termfix)
            for (int i = 0; i < NX; ++i)
                if ((fN >> i) & 1) bidx[nb++] = i;
        nact = (N + 1) * NZ - NU - nf0 - nfN;
        for (int i = 0; i < NQ; ++i) wcost[i] = (FAM == VBOC_FAMILY_VBOC) ? pb.p[i] : 0.0;
        for (int sc = 0; sc < 3; ++sc) {
            const double *l = sc == 0 ? pb.lbx0 : (sc == 1 ? pb.lbx : pb.lbxN);
            const double *u = sc == 0 ? pb.ubx0 : (sc == 1 ? pb.ubx : pb.ubxN);
            for (int i = 0; i < NZ; ++i) {
                g(L::LB + sc * NZ + i) = i < NU ? pb.lbu[i] : l[i - NU];
                g(L::UB + sc * NZ + i) = i < NU ? pb.ubu[i] : u[i - NU];
            }
        }
        for (int k = 0; k <= N; ++k) {
            for (int i = 0; i < NZ; ++i)
                s(k, L::Z + i) = i < NU ? (k < N ? pb.ug[k * NU + i] : 0.0) : pb.xg[(size_t)k * pb.nxr + i - NU];
            for (int i = 0; i < NX; ++i) s(k, L::PI + i) = 0.0;
            for (int c = 0; c < NC; ++c) s(k, L::LAM + c) = 0.0;
        }
    }
    VB_HD void store_solution(const Prob &pb, const vboc_stats &st) const {
        for (int k = 0; k <= N; ++k) {
            for (int i = 0; i < NX; ++i) pb.x[(size_t)k * pb.nxr + i] = s(k, L::Z + NU + i);
            if (pb.nxr > NX) pb.x[(size_t)k * pb.nxr + NX] = h;
            if (k < N)
                for (int i = 0; i < NU; ++i) pb.u[k * NU + i] = s(k, L::Z + i);
        }
        *pb.st = st;
    }

    // ---------------------------------------------------------------- linearisation
    VB_HD void linearize() {
#pragma unroll 1
        for (int k = 0; k < N; ++k) {
            double x[NX], u[NU], xn[NX], Phi[NX][NZ];
            for (int i = 0; i < NU; ++i) u[i] = s(k, L::Z + i);
            for (int i = 0; i < NX; ++i) x[i] = s(k, L::Z + NU + i);
            if constexpr (DTS)
                rk4_sens_dts<NQ>(x, u, xn, Phi);
            else
                rk4_sens<NQ>(x, u, h, xn, Phi);
            for (int i = 0; i < NX; ++i) {
                for (int j = 0; j < NZ; ++j) s(k, L::BA + i * NZ + j) = Phi[i][j];
                s(k, L::BD + i) = xn[i] - s(k + 1, L::Z + NU + i);
            }
        }
    }

    // ---------------------------------------------------------------- NLP residuals
    VB_HD bool nlp_residuals(double &rs, double &re, double &ri, double &rc) {
        double vs = 0, ve = 0, vi = 0, vc = 0;
        bool nan = false;
#pragma unroll 1
        for (int k = 0; k <= N; ++k) {
            double r[NZ];
            for (int i = 0; i < NZ; ++i) {
                r[i] = 0.0;
                if (k == N && i < NU) continue;
                double z = s(k, L::Z + i);
                double a = cost_g(k, i, z);
                if (k < N)
                    for (int m = 0; m < NX; ++m) a += s(k, L::BA + m * NZ + i) * s(k, L::PI + m);
                if (k > 0 && i >= NU) a -= s(k - 1, L::PI + i - NU);
                if (active(k, i)) {
                    double ll = s(k, L::LAM + i), lu = s(k, L::LAM + NZ + i);
                    double fl = lb(k, i) - z, fu = z - ub(k, i);
                    a += lu - ll;
                    vi = fmax(vi, fmax(fl, fu));
                    vc = fmax(vc, fmax(fabs(ll * fl), fabs(lu * fu)));
                } else if (k == N) {
                    a = 0.0;
                    vi = fmax(vi, fabs(z - g(L::CN + i - NU)));
                }
                r[i] = a;
            }
            if (k == 0) {
                double x0[NX], e[NX];
                proj0(r + NU);
                for (int i = 0; i < NX; ++i) x0[i] = s(0, L::Z + NU + i);
                eq0_violation(x0, e);
                for (int i = 0; i < NX; ++i) vi = fmax(vi, fabs(e[i]));
            }
            for (int i = 0; i < NZ; ++i) nan |= (r[i] != r[i]), vs = fmax(vs, fabs(r[i]));
            if (k < N)
                for (int i = 0; i < NX; ++i) {
                    double v = s(k, L::BD + i);
                    nan |= (v != v), ve = fmax(ve, fabs(v));
                }
        }
        rs = vs, re = ve, ri = vi, rc = vc;
        return !nan;
    }

    // ---------------------------------------------------------------- QP: cold start
    VB_HD void qp_init() {
        const double thr0 = 0.1;
#pragma unroll 1
        for (int k = 0; k <= N; ++k) {
            for (int i = 0; i < NZ; ++i) {
                double z = s(k, L::Z + i), v = 0.0, tl = 0, tu = 0, ll = 0, lu = 0;
                if (active(k, i)) {
                    double lbd = lb(k, i) - z, ubd = ub(k, i) - z;
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
                    v = g(L::CN + i - NU) - z;
                }
                s(k, L::DZ + i) = v;
                s(k, L::LAMQ + i) = ll, s(k, L::LAMQ + NZ + i) = lu;
                s(k, L::TQ + i) = tl, s(k, L::TQ + NZ + i) = tu;
            }
            if (k < N)
                for (int i = 0; i < NX; ++i) s(k, L::PIQ + i) = 0.0;
        }
        double x0[NX], e[NX];
        for (int i = 0; i < NX; ++i) x0[i] = s(0, L::Z + NU + i) + s(0, L::DZ + NU + i);
        eq0_violation(x0, e);
        for (int i = 0; i < NX; ++i) s(0, L::DZ + NU + i) -= e[i];
    }

    // ---------------------------------------------------------------- QP: residuals
    // writes RB, HH, RR; returns mu and the four norms
    VB_HD double qp_residuals(double &ng, double &nb_, double &nd, double &nm, bool &nan) {
        double vg = 0, vb = 0, vd = 0, vm = 0, mu = 0;
        bool bad = false;
#pragma unroll 1
        for (int k = 0; k <= N; ++k) {
            double r[NZ], bar[NZ];
            for (int i = 0; i < NZ; ++i) {
                double v = s(k, L::DZ + i), z = s(k, L::Z + i);
                double hd = cost_h(k, i), hh = hd + o.qp_reg_prim, b = 0.0;
                double a = hd * v + cost_g(k, i, z);
                if (k < N)
                    for (int m = 0; m < NX; ++m) a += s(k, L::BA + m * NZ + i) * s(k, L::PIQ + m);
                if (k > 0 && i >= NU) a -= s(k - 1, L::PIQ + i - NU);
                if (active(k, i)) {
                    double ll = s(k, L::LAMQ + i), lu = s(k, L::LAMQ + NZ + i);
                    double tl = s(k, L::TQ + i), tu = s(k, L::TQ + NZ + i);
                    a += lu - ll;
                    double dl = (lb(k, i) - z) - v + tl, du = v - (ub(k, i) - z) + tu;
                    double ml = ll * tl, mu_ = lu * tu;
                    bad |= (dl != dl) | (du != du) | (ml != ml) | (mu_ != mu_);
                    vd = fmax(vd, fmax(fabs(dl), fabs(du)));
                    vm = fmax(vm, fmax(fabs(ml), fabs(mu_)));
                    mu += ml + mu_;
                    double itl = 1.0 / tl, itu = 1.0 / tu;
                    hh += ll * itl + lu * itu;
                    b = (ml - ll * dl) * itl - (mu_ - lu * du) * itu;
                } else if (k == N) {
                    a = 0.0;
                }
                r[i] = a, bar[i] = b;
                s(k, L::HH + i) = hh;
            }
            if (k == 0) proj0(r + NU);
            for (int i = 0; i < NZ; ++i) {
                bad |= (r[i] != r[i]);
                vg = fmax(vg, fabs(r[i]));
                s(k, L::RR + i) = r[i] + bar[i];
            }
            if (k < N)
                for (int i = 0; i < NX; ++i) {
                    double a = s(k, L::BD + i) - s(k + 1, L::DZ + NU + i);
                    for (int j = 0; j < NZ; ++j) a += s(k, L::BA + i * NZ + j) * s(k, L::DZ + j);
                    s(k, L::RB + i) = a;
                    bad |= (a != a);
                    vb = fmax(vb, fabs(a));
                }
        }
        double x0[NX], e[NX];
        for (int i = 0; i < NX; ++i) x0[i] = s(0, L::Z + NU + i) + s(0, L::DZ + NU + i);
        eq0_violation(x0, e);
        for (int i = 0; i < NX; ++i) {
            double en = ((fixedN >> i) & 1) ? s(N, L::Z + NU + i) + s(N, L::DZ + NU + i) - g(L::CN + i) : 0.0;
            g(L::E0 + i) = e[i], g(L::EN + i) = en;
            vb = fmax(vb, fmax(fabs(e[i]), fabs(en)));
        }
        ng = vg, nb_ = vb, nd = vd, nm = vm, nan = bad;
        return nact ? mu / (2.0 * nact) : 0.0;
    }

    VB_HD double rhs_of(int k, int i, int mode, double sm) const {
        if (mode == 3) return 0.0;
        if (mode == 0) return s(k, L::RR + i);
        return (mode == 1 ? s(k, L::Q1 + i) : s(k, L::RR + i)) - sm * s(k, L::Q2 + i);
    }

    // ---------------------------------------------------------------- Riccati: backward sweep
    // mode 0 factorise + predictor rhs, 1 corrector, 2 centering only, 3 border basis solve for the
    // fixed terminal component bidx[bj] (unit terminal gradient, all other right-hand sides zero)
    VB_HD bool backward(int mode, double sm, double *dx0, int bj = 0) {
        const bool factor = mode == 0;
        double P[NX][NX], pv[NX];
        for (int i = 0; i < NX; ++i) {
            bool fx = termfix && ((fixedN >> i) & 1);
            double hh = s(N, L::HH + NU + i), rr = rhs_of(N, NU + i, mode, sm);
            if (mode == 3) rr = (i == bidx[bj]) ? 1.0 : 0.0;
            g(L::HHN + i) = fx ? 0.0 : hh;
            g(L::RN + i) = fx ? 0.0 : rr;
            pv[i] = fx ? 0.0 : rr;
            for (int j = 0; j < NX; ++j) P[i][j] = (j == i && !fx) ? hh : 0.0;
        }
        bool ok = true;
#pragma unroll 1
        for (int k = N - 1; k >= 0; --k) {
            const bool last = (k == N - 1) && termfix;
            double BA[NX][NZ], m[NZ];
            for (int i = 0; i < NX; ++i)
                for (int j = 0; j < NZ; ++j) BA[i][j] = s(k, L::BA + i * NZ + j);
            double M[NZ][NZ];
            if (factor) {
                double T[NX][NZ], tb[NX];  // P+ [B A], P+ beta
                for (int i = 0; i < NX; ++i) {
                    double a = 0.0;
                    for (int mm = 0; mm < NX; ++mm) a += P[i][mm] * s(k, L::RB + mm);
                    tb[i] = a;
                    for (int j = 0; j < NZ; ++j) {
                        double t = 0.0;
                        for (int mm = 0; mm < NX; ++mm) t += P[i][mm] * BA[mm][j];
                        T[i][j] = t;
                    }
                }
                for (int a_ = 0; a_ < NZ; ++a_) {
                    for (int b_ = 0; b_ <= a_; ++b_) {
                        double t = (a_ == b_) ? s(k, L::HH + a_) : 0.0;
                        for (int mm = 0; mm < NX; ++mm) t += BA[mm][a_] * T[mm][b_];
                        M[a_][b_] = t, M[b_][a_] = t;
                    }
                    double mb = 0.0;
                    for (int i = 0; i < NX; ++i) mb += BA[i][a_] * tb[i];
                    s(k, L::MB + a_) = mb;
                    m[a_] = mb;
                }
            } else {
                for (int a_ = 0; a_ < NZ; ++a_) m[a_] = mode == 3 ? 0.0 : s(k, L::MB + a_);
                if (last)
                    for (int a_ = 0; a_ < NZ; ++a_)
                        for (int b_ = 0; b_ < NZ; ++b_) M[a_][b_] = g(L::MF + a_ * NZ + b_);
            }
            for (int a_ = 0; a_ < NZ; ++a_) {
                double t = m[a_] + rhs_of(k, a_, mode, sm);
                for (int i = 0; i < NX; ++i) t += BA[i][a_] * pv[i];
                m[a_] = t;
            }
            if (!last) {
                double Lu[NU][NU], di[NU], y[NU], Lxu[NX][NU];
                if (factor) {
                    for (int j = 0; j < NU; ++j) {
                        double d = M[j][j];
                        for (int c = 0; c < j; ++c) d -= Lu[j][c] * Lu[j][c];
                        di[j] = d > 0.0 ? VB_RSQRT(d) : 0.0;
                        for (int i = j + 1; i < NU; ++i) {
                            double a = M[i][j];
                            for (int c = 0; c < j; ++c) a -= Lu[i][c] * Lu[j][c];
                            Lu[i][j] = a * di[j];
                        }
                    }
                    for (int i = 0; i < NX; ++i)
                        for (int c = 0; c < NU; ++c) {
                            double a = M[NU + i][c];
                            for (int c2 = 0; c2 < c; ++c2) a -= Lxu[i][c2] * Lu[c][c2];
                            Lxu[i][c] = a * di[c];
                            s(k, L::LXU + i * NU + c) = Lxu[i][c];
                        }
                    for (int i = 0; i < NU; ++i)
                        for (int c = 0; c <= i; ++c) s(k, L::LUU + i * NU + c) = (c == i) ? di[i] : Lu[i][c];
                    for (int i = 0; i < NX; ++i)
                        for (int j = 0; j <= i; ++j) {
                            double a = M[NU + i][NU + j];
                            for (int c = 0; c < NU; ++c) a -= Lxu[i][c] * Lxu[j][c];
                            P[i][j] = a, P[j][i] = a;
                        }
                    for (int i = 0; i < NX; ++i)
                        for (int j = 0; j < NX; ++j) s(k, L::P + i * NX + j) = P[i][j];
                } else {
                    for (int i = 0; i < NU; ++i) {
                        di[i] = s(k, L::LUU + i * NU + i);
                        for (int c = 0; c < i; ++c) Lu[i][c] = s(k, L::LUU + i * NU + c);
                    }
                    for (int i = 0; i < NX; ++i)
                        for (int c = 0; c < NU; ++c) Lxu[i][c] = s(k, L::LXU + i * NU + c);
                }
                for (int c = 0; c < NU; ++c) {
                    double a = m[c];
                    for (int c2 = 0; c2 < c; ++c2) a -= Lu[c][c2] * y[c2];
                    y[c] = a * di[c];
                    s(k, L::YV + c) = y[c];
                }
                for (int i = 0; i < NX; ++i) {
                    double p = m[NU + i];
                    for (int c = 0; c < NU; ++c) p -= Lxu[i][c] * y[c];
                    pv[i] = p;
                    s(k, L::PV + i) = p;
                }
            } else {
                // terminal velocity equalities through the last control: du = K dx + k0
                double Gi[NU][NU], K[NU][NX], k0[NU], tmp[NZ];
                if (factor) {
                    double G[NU][NU];
                    for (int a = 0; a < NU; ++a)
                        for (int b = 0; b < NU; ++b) G[a][b] = BA[NQ + a][b];
                    ok = inverse_small(G, Gi) && ok;
                    for (int a = 0; a < NU; ++a)
                        for (int j = 0; j < NX; ++j) {
                            double v = 0.0;
                            for (int b = 0; b < NU; ++b) v -= Gi[a][b] * BA[NQ + b][NU + j];
                            K[a][j] = v;
                            s(k, L::LXU + j * NU + a) = v;
                        }
                    for (int a = 0; a < NU; ++a)
                        for (int b = 0; b < NU; ++b) s(k, L::LUU + a * NU + b) = Gi[a][b];
                    for (int a_ = 0; a_ < NZ; ++a_)
                        for (int b_ = 0; b_ < NZ; ++b_) g(L::MF + a_ * NZ + b_) = M[a_][b_];
                } else {
                    for (int a = 0; a < NU; ++a) {
                        for (int b = 0; b < NU; ++b) Gi[a][b] = s(k, L::LUU + a * NU + b);
                        for (int j = 0; j < NX; ++j) K[a][j] = s(k, L::LXU + j * NU + a);
                    }
                }
                for (int a = 0; a < NU; ++a) {
                    double v = 0.0;
                    for (int b = 0; b < NU; ++b) v -= Gi[a][b] * (g(L::EN + NQ + b) + s(k, L::RB + NQ + b));
                    k0[a] = v;
                    g(L::K0 + a) = v;
                }
                for (int i = 0; i < NZ; ++i) {
                    double v = m[i];
                    for (int a = 0; a < NU; ++a) v += M[i][a] * k0[a];
                    tmp[i] = v;
                    g(L::ML + i) = m[i];
                }
                if (factor) {
                    for (int i = 0; i < NX; ++i)
                        for (int j = 0; j <= i; ++j) {
                            double a = M[NU + i][NU + j];
                            for (int c = 0; c < NU; ++c) {
                                a += K[c][i] * M[c][NU + j] + M[NU + i][c] * K[c][j];
                                for (int c2 = 0; c2 < NU; ++c2) a += K[c][i] * M[c][c2] * K[c2][j];
                            }
                            P[i][j] = a, P[j][i] = a;
                        }
                    for (int i = 0; i < NX; ++i)
                        for (int j = 0; j < NX; ++j) s(k, L::P + i * NX + j) = P[i][j];
                }
                for (int j = 0; j < NX; ++j) {
                    double p = tmp[NU + j];
                    for (int a = 0; a < NU; ++a) p += K[a][j] * tmp[a];
                    pv[j] = p;
                    s(k, L::PV + j) = p;
                }
            }
        }
        // stage 0: dx0 = -e0 + Z0 dy, (Z0'P0 Z0) dy = -Z0'(p0 - P0 e0)
        if (factor) {
            double T[NX][NX], Lz[NX][NX], dzi[NX];
            for (int i = 0; i < NX; ++i) {
                double a = 0.0;
                for (int j = 0; j < NX; ++j) a += P[i][j] * g(L::E0 + j);
                g(L::PE + i) = a;
                for (int c = 0; c < NX; ++c) {
                    double t = 0.0;
                    for (int j = 0; j < NX; ++j) t += P[i][j] * g(L::Z0 + j * NX + c);
                    T[i][c] = t;
                }
            }
            for (int a_ = 0; a_ < NX; ++a_)
                for (int b_ = 0; b_ <= a_; ++b_) {
                    double t = 0.0;
                    for (int i = 0; i < NX; ++i) t += g(L::Z0 + i * NX + a_) * T[i][b_];
                    Lz[a_][b_] = t;
                }
            for (int j = 0; j < NX; ++j) {
                double d = Lz[j][j];
                for (int c = 0; c < j; ++c) d -= Lz[j][c] * Lz[j][c];
                dzi[j] = d > 0.0 ? VB_RSQRT(d) : 0.0;
                for (int i = j + 1; i < NX; ++i) {
                    double a = Lz[i][j];
                    for (int c = 0; c < j; ++c) a -= Lz[i][c] * Lz[j][c];
                    Lz[i][j] = a * dzi[j];
                }
            }
            for (int i = 0; i < NX; ++i) {
                g(L::DZI + i) = dzi[i];
                for (int c = 0; c < i; ++c) g(L::LZ + i * NX + c) = Lz[i][c];
            }
        }
        {
            double y[NX], dy[NX];
            const double e0s = mode == 3 ? 0.0 : 1.0;  // the basis solves are homogeneous
            for (int i = 0; i < NX; ++i) {
                double a = 0.0;
                for (int mm = 0; mm < NX; ++mm) a -= g(L::Z0 + mm * NX + i) * (pv[mm] - e0s * g(L::PE + mm));
                for (int j = 0; j < i; ++j) a -= g(L::LZ + i * NX + j) * y[j];
                y[i] = a * g(L::DZI + i);
            }
            for (int i = NX - 1; i >= 0; --i) {
                double a = y[i];
                for (int j = i + 1; j < NX; ++j) a -= g(L::LZ + j * NX + i) * dy[j];
                dy[i] = a * g(L::DZI + i);
            }
            for (int i = 0; i < NX; ++i) {
                double a = -e0s * g(L::E0 + i);
                for (int c = 0; c < NX; ++c) a += g(L::Z0 + i * NX + c) * dy[c];
                dx0[i] = a;
            }
        }
        return ok;
    }

    VB_HD bool inverse_small(const double (&G)[NU][NU], double (&Gi)[NU][NU]) const {
        if constexpr (NU == 1) {
            Gi[0][0] = 1.0 / G[0][0];
            return G[0][0] != 0.0 && G[0][0] == G[0][0];
        } else if constexpr (NU == 2) {
            double det = G[0][0] * G[1][1] - G[0][1] * G[1][0], r = 1.0 / det;
            Gi[0][0] = G[1][1] * r, Gi[0][1] = -G[0][1] * r, Gi[1][0] = -G[1][0] * r, Gi[1][1] = G[0][0] * r;
            return det != 0.0 && det == det;
        } else {
            double c00 = G[1][1] * G[2][2] - G[1][2] * G[2][1], c01 = G[1][2] * G[2][0] - G[1][0] * G[2][2];
            double c02 = G[1][0] * G[2][1] - G[1][1] * G[2][0];
            double det = G[0][0] * c00 + G[0][1] * c01 + G[0][2] * c02, r = 1.0 / det;
            Gi[0][0] = c00 * r, Gi[1][0] = c01 * r, Gi[2][0] = c02 * r;
            Gi[0][1] = (G[0][2] * G[2][1] - G[0][1] * G[2][2]) * r;
            Gi[1][1] = (G[0][0] * G[2][2] - G[0][2] * G[2][0]) * r;
            Gi[2][1] = (G[0][1] * G[2][0] - G[0][0] * G[2][1]) * r;
            Gi[0][2] = (G[0][1] * G[1][2] - G[0][2] * G[1][1]) * r;
            Gi[1][2] = (G[0][2] * G[1][0] - G[0][0] * G[1][2]) * r;
            Gi[2][2] = (G[0][0] * G[1][1] - G[0][1] * G[1][0]) * r;
            return det != 0.0 && det == det;
        }
    }

    // ---------------------------------------------------------------- Riccati: forward sweep + constraint steps
    // DV, DPI, the slack / multiplier steps, the maximum step to the boundary and the sums of
    // mu(alpha) = (S0 + alpha S1 + alpha^2 S2) / nc.  mode 0 stores the second-order products and the
    // corrector gradient pieces Q1, Q2; modes 1/2 store DT, DLAM.
    // slack / multiplier steps of the bound constraints of stage k for the primal step dz: ratio test and
    // the sums of mu(alpha); mode 0 stores the second-order products and Q1, Q2, modes 1/2 store DT, DLAM
    VB_HD void con_stage(int k, const double *dz, int mode, double sm, double &al, double &s0, double &s1,
                         double &s2) {
            for (int i = 0; i < NZ; ++i) {
                double q1 = 0.0, q2 = 0.0;
                if (active(k, i)) {
                    double z = s(k, L::Z + i), v = s(k, L::DZ + i);
#pragma unroll
                    for (int sd = 0; sd < 2; ++sd) {
                        double lam = s(k, L::LAMQ + sd * NZ + i), t = s(k, L::TQ + sd * NZ + i);
                        double rd = sd ? v - (ub(k, i) - z) + t : (lb(k, i) - z) - v + t;
                        double rm = lam * t;
                        if (mode == 1) rm += s(k, L::PROD + sd * NZ + i) - sm;
                        if (mode == 2) rm -= sm;
                        double dtt = (sd ? -dz[i] : dz[i]) - rd;
                        double it = 1.0 / t;
                        double dl = -(rm + lam * dtt) * it;
                        if (dtt < 0.0 && t + al * dtt < 0.0) al = fmin(al, -t / dtt);
                        if (dl < 0.0 && lam + al * dl < 0.0) al = fmin(al, -lam / dl);
                        s0 += lam * t, s1 += lam * dtt + t * dl, s2 += dtt * dl;
                        if (mode == 0) {
                            double pr = dtt * dl;
                            s(k, L::PROD + sd * NZ + i) = pr;
                            q1 += sd ? -pr * it : pr * it;
                            q2 += sd ? -it : it;
                        } else {
                            s(k, L::DT + sd * NZ + i) = dtt, s(k, L::DLAM + sd * NZ + i) = dl;
                        }
                    }
                }
                if (mode == 0) s(k, L::Q1 + i) = s(k, L::RR + i) + q1, s(k, L::Q2 + i) = q2;
            }
    }
    VB_HD double con_pass(int mode, double sm, double &S0, double &S1, double &S2) {
        double al = 1.0, s0 = 0, s1 = 0, s2 = 0;
#pragma unroll 1
        for (int k = 0; k <= N; ++k) {
            double dz[NZ];
            for (int i = 0; i < NZ; ++i) dz[i] = s(k, L::DV + i);
            con_stage(k, dz, mode, sm, al, s0, s1, s2);
        }
        S0 = s0, S1 = s1, S2 = s2;
        return al;
    }

    // border mode: terminal multipliers from the basis solves, then DV, DPI of the full step
    VB_HD bool border_combine() {
        double Sm[NX][NX + 1], nuv[NX];
        for (int i = 0; i < nb; ++i) {
            for (int j = 0; j < nb; ++j) Sm[i][j] = s(N, L::DVB + (size_t)j * NZ + NU + bidx[i]);
            Sm[i][nb] = -(g(L::EN + bidx[i]) + s(N, L::DV + NU + bidx[i]));
        }
        for (int c = 0; c < nb; ++c) {  // Gaussian elimination with partial pivoting
            int p = c;
            for (int i = c + 1; i < nb; ++i)
                if (fabs(Sm[i][c]) > fabs(Sm[p][c])) p = i;
            if (Sm[p][c] == 0.0 || Sm[p][c] != Sm[p][c]) return false;
            for (int j = 0; j <= nb; ++j) {
                double t = Sm[c][j];
                Sm[c][j] = Sm[p][j], Sm[p][j] = t;
            }
            for (int i = c + 1; i < nb; ++i) {
                double f = Sm[i][c] / Sm[c][c];
                for (int j = c; j <= nb; ++j) Sm[i][j] -= f * Sm[c][j];
            }
        }
        for (int i = nb - 1; i >= 0; --i) {
            double a = Sm[i][nb];
            for (int j = i + 1; j < nb; ++j) a -= Sm[i][j] * nuv[j];
            nuv[i] = a / Sm[i][i];
        }
#pragma unroll 1
        for (int k = 0; k <= N; ++k)
            for (int j = 0; j < nb; ++j) {
                for (int i = 0; i < NZ; ++i) s(k, L::DV + i) += nuv[j] * s(k, L::DVB + (size_t)j * NZ + i);
                if (k < N)
                    for (int i = 0; i < NX; ++i) s(k, L::DPI + i) += nuv[j] * s(k, L::DPIB + (size_t)j * NX + i);
            }
        return true;
    }

    // raw: only the recurrence (DV, DPI, or the basis arrays DVB / DPIB when mode == 3); the constraint
    // steps are then done by con_pass after the border combination.
    VB_HD double forward(int mode, double sm, const double *dx0, double &S0, double &S1, double &S2, bool raw = false,
                         int bj = 0) {
        const size_t fDV = mode == 3 ? L::DVB + (size_t)bj * NZ : L::DV;
        const size_t fDPI = mode == 3 ? L::DPIB + (size_t)bj * NX : L::DPI;
        double dx[NX], al = 1.0, s0 = 0, s1 = 0, s2 = 0;
        for (int i = 0; i < NX; ++i) dx[i] = dx0[i];
#pragma unroll 1
        for (int k = 0; k <= N; ++k) {
            const bool last = (k == N - 1) && termfix;
            double dz[NZ];
            for (int i = 0; i < NX; ++i) dz[NU + i] = dx[i];
            for (int c = 0; c < NU; ++c) dz[c] = 0.0;
            if (k < N) {
                if (!last) {
                    double t[NU];
                    for (int c = 0; c < NU; ++c) {
                        double a = s(k, L::YV + c);
                        for (int j = 0; j < NX; ++j) a += s(k, L::LXU + j * NU + c) * dx[j];
                        t[c] = a;
                    }
                    for (int c = NU - 1; c >= 0; --c) {
                        double a = -t[c];
                        for (int c2 = c + 1; c2 < NU; ++c2) a -= s(k, L::LUU + c2 * NU + c) * dz[c2];
                        dz[c] = a * s(k, L::LUU + c * NU + c);
                    }
                } else {
                    for (int a_ = 0; a_ < NU; ++a_) {
                        double a = g(L::K0 + a_);
                        for (int j = 0; j < NX; ++j) a += s(k, L::LXU + j * NU + a_) * dx[j];
                        dz[a_] = a;
                    }
                }
            }
            for (int i = 0; i < NZ; ++i) s(k, fDV + i) = dz[i];
            if (!raw) con_stage(k, dz, mode, sm, al, s0, s1, s2);
            if (k < N) {
                double dxn[NX];
                for (int i = 0; i < NX; ++i) {
                    double a = mode == 3 ? 0.0 : s(k, L::RB + i);
                    for (int j = 0; j < NZ; ++j) a += s(k, L::BA + i * NZ + j) * dz[j];
                    dxn[i] = a;
                }
                // dpi_k = P_{k+1} dx_{k+1} + p_{k+1}
                if (k + 1 < N) {
                    for (int i = 0; i < NX; ++i) {
                        double a = s(k + 1, L::PV + i);
                        for (int j = 0; j < NX; ++j) a += s(k + 1, L::P + i * NX + j) * dxn[j];
                        s(k, fDPI + i) = a;
                    }
                } else {
                    double nuv[NU];
                    if (last) {
                        double tu[NU];
                        for (int a_ = 0; a_ < NU; ++a_) {
                            double a = g(L::ML + a_);
                            for (int j = 0; j < NZ; ++j) a += g(L::MF + a_ * NZ + j) * dz[j];
                            tu[a_] = a;
                        }
                        for (int b = 0; b < NU; ++b) {
                            double a = 0.0;
                            for (int a_ = 0; a_ < NU; ++a_) a -= s(k, L::LUU + a_ * NU + b) * tu[a_];
                            nuv[b] = a;
                        }
                    }
                    for (int i = 0; i < NX; ++i) {
                        double a = g(L::HHN + i) * dxn[i] + g(L::RN + i);
                        if (last && i >= NQ) a = nuv[i - NQ];
                        s(k, fDPI + i) = a;
                    }
                }
                for (int i = 0; i < NX; ++i) dx[i] = dxn[i];
            }
        }
        S0 = s0, S1 = s1, S2 = s2;
        return al;
    }

    // ---------------------------------------------------------------- IPM
    VB_HD int ipm_solve(int &iters) {
        qp_init();
        double rg = 0, rb = 0, rd = 0, rm = 0, alpha = 1.0, mu = 0.0;
        bool nan = false, ok = true;
        const double nc = 2.0 * nact;
        int kk = 0;
#pragma unroll 1
        for (;; ++kk) {
            mu = qp_residuals(rg, rb, rd, rm, nan);
            if (!(kk < o.qp_iter_max && alpha > o.qp_alpha_min && !nan &&
                  (rg > o.qp_tol_stat || rb > o.qp_tol_eq || rd > o.qp_tol_ineq || rm > o.qp_tol_comp)))
                break;
            double sm = 0.0, m_aff = 0.0;
#pragma unroll 1
            for (int ph = 0; ph < 3; ++ph) {
                double S0, S1, S2, dx0[NX];
                ok = backward(ph, sm, dx0) && ok;
                if (!ok) break;
                if (nb == 0) {
                    alpha = forward(ph, sm, dx0, S0, S1, S2);
                } else {
                    forward(ph, sm, dx0, S0, S1, S2, true);
                    if (ph == 0)
                        for (int j = 0; j < nb; ++j) {
                            double dxb[NX], d0, d1, d2;
                            backward(3, 0.0, dxb, j);
                            forward(3, 0.0, dxb, d0, d1, d2, true, j);
                        }
                    ok = border_combine() && ok;
                    if (!ok) break;
                    alpha = con_pass(ph, sm, S0, S1, S2);
                }
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
#pragma unroll 1
            for (int k = 0; k <= N; ++k) {
                for (int i = 0; i < NZ; ++i) s(k, L::DZ + i) += as * s(k, L::DV + i);
                if (k < N)
                    for (int i = 0; i < NX; ++i) s(k, L::PIQ + i) += as * s(k, L::DPI + i);
                for (int c = 0; c < NC; ++c) {
                    int i = c >= NZ ? c - NZ : c;
                    if (active(k, i)) {
                        s(k, L::LAMQ + c) = fmax(s(k, L::LAMQ + c) + as * s(k, L::DLAM + c), o.qp_lam_min);
                        s(k, L::TQ + c) = fmax(s(k, L::TQ + c) + as * s(k, L::DT + c), o.qp_t_min);
                    }
                }
            }
        }
        iters = kk;
        {
            double r[NX], rp[NX];
            for (int i = 0; i < NX; ++i) {
                double a = cost_h(0, NU + i) * s(0, L::DZ + NU + i) + cost_g(0, NU + i, s(0, L::Z + NU + i));
                for (int m = 0; m < NX; ++m) a += s(0, L::BA + m * NZ + NU + i) * s(0, L::PIQ + m);
                if (active(0, NU + i)) a += s(0, L::LAMQ + NZ + NU + i) - s(0, L::LAMQ + NU + i);
                r[i] = rp[i] = a;
                g(L::NUNQ + i) = ((fixedN >> i) & 1)
                                     ? s(N - 1, L::PIQ + i) - cost_g(N, NU + i, s(N, L::Z + NU + i)) -
                                           cost_h(N, NU + i) * s(N, L::DZ + NU + i)
                                     : 0.0;
            }
            proj0(rp);
            for (int i = 0; i < NX; ++i) g(L::NU0Q + i) = r[i] - rp[i];
        }
        if (!ok || nan || mu != mu) return 3;
        if (rg > o.qp_tol_stat || rb > o.qp_tol_eq || rd > o.qp_tol_ineq || rm > o.qp_tol_comp)
            return kk >= o.qp_iter_max ? 1 : 2;
        return 0;
    }

    // ---------------------------------------------------------------- merit function / line search
    VB_HD double total_cost(size_t zf) const {
        if (FAM == VBOC_FAMILY_VBOC) {
            double c = wtdt;
            for (int i = 0; i < NQ; ++i) c += wcost[i] * s(0, zf + NU + NQ + i);
            if (DTS)
                for (int k = 0; k < N; ++k) c += wt * s(k, zf + NU + 2 * NQ);
            return c;
        }
        double a = 0.0;
        for (int k = 0; k <= N; ++k) {
            double q = 0.0;
            for (int i = 0; i < NQ; ++i) q += s(k, zf + NU + NQ + i) * s(k, zf + NU + NQ + i);
            a += (k < N ? h : 1.0) * q;
        }
        return a;
    }
    VB_HD double merit(size_t zf) const {
        double m = 0.0;
#pragma unroll 1
        for (int k = 0; k <= N; ++k) {
            if (k < N) {
                double x[NX], u[NU], xn[NX];
                for (int i = 0; i < NU; ++i) u[i] = s(k, zf + i);
                for (int i = 0; i < NX; ++i) x[i] = s(k, zf + NU + i);
                if constexpr (DTS)
                    rk4_step_dts<NQ>(x, u, xn);
                else
                    rk4_step<NQ, double>(x, u, h, xn);
                for (int i = 0; i < NX; ++i) m += s(k, L::WDYN + i) * fabs(xn[i] - s(k + 1, zf + NU + i));
            }
            for (int i = 0; i < NZ; ++i)
                if (active(k, i)) {
                    double z = s(k, zf + i), fl = lb(k, i) - z, fu = z - ub(k, i);
                    if (fl > 0) m += s(k, L::WB + i) * fl;
                    if (fu > 0) m += s(k, L::WB + NZ + i) * fu;
                }
        }
        m += total_cost(zf);
        double x0[NX], e[NX];
        for (int i = 0; i < NX; ++i) x0[i] = s(0, zf + NU + i);
        eq0_violation(x0, e);
        for (int i = 0; i < NX; ++i) {
            m += g(L::W0 + i) * fabs(e[i]);
            if ((fixedN >> i) & 1) m += g(L::WN + i) * fabs(s(N, zf + NU + i) - g(L::CN + i));
        }
        return m;
    }
    VB_HD double line_search(int sqp_iter, int &evals) {
#pragma unroll 1
        for (int k = 0; k <= N; ++k) {
            if (k < N)
                for (int i = 0; i < NX; ++i) {
                    double a = fabs(s(k, L::PIQ + i));
                    s(k, L::WDYN + i) = sqp_iter == 0 ? a : fmax(a, 0.5 * (s(k, L::WDYN + i) + a));
                }
            for (int c = 0; c < NC; ++c) {
                double a = fabs(s(k, L::LAMQ + c));
                s(k, L::WB + c) = sqp_iter == 0 ? a : fmax(a, 0.5 * (s(k, L::WB + c) + a));
            }
        }
        for (int i = 0; i < NX; ++i) {
            double a = fabs(g(L::NU0Q + i)), b = fabs(g(L::NUNQ + i));
            g(L::W0 + i) = sqp_iter == 0 ? a : fmax(a, 0.5 * (g(L::W0 + i) + a));
            g(L::WN + i) = sqp_iter == 0 ? b : fmax(b, 0.5 * (g(L::WN + i) + b));
        }
        double m0 = merit(L::Z), alpha = 1.0;
#pragma unroll 1
        for (;;) {
            for (int k = 0; k <= N; ++k)
                for (int i = 0; i < NZ; ++i) s(k, L::ZT + i) = s(k, L::Z + i) + alpha * s(k, L::DZ + i);
            double m1 = merit(L::ZT);
            ++evals;
            if (m1 < m0) break;
            if (alpha * o.alpha_reduction < o.alpha_min) break;
            alpha *= o.alpha_reduction;
        }
        return alpha;
    }

    // ---------------------------------------------------------------- one SQP iteration
    // Returns true when the problem is finished (status in st).
    VB_HD bool sqp_iteration(LaneState &ls, int mode) {
        vboc_stats &st = ls.st;
        const int maxit = mode == VBOC_MODE_RTI ? 1 : o.max_iter;
        linearize();
        bool finite = nlp_residuals(st.res_stat, st.res_eq, st.res_ineq, st.res_comp);
        if (mode == VBOC_MODE_SQP || ls.it > 0) {
            if (!finite) {
                st.status = VBOC_FAILURE;
                return true;
            }
            if (mode == VBOC_MODE_SQP && st.res_stat < o.tol_stat && st.res_eq < o.tol_eq &&
                st.res_ineq < o.tol_ineq && st.res_comp < o.tol_comp) {
                st.status = VBOC_SUCCESS;
                return true;
            }
        }
        if (ls.it >= maxit) {
            st.status = mode == VBOC_MODE_RTI ? VBOC_SUCCESS : VBOC_MAXITER;
            return true;
        }
        int qit = 0;
        int qs = ipm_solve(qit);
        st.qp_iter += qit, st.qp_status = qs, st.sqp_iter = ls.it + 1;
        if (qs != 0 && qs != 1) {
            st.status = VBOC_QP_FAILURE;
            return true;
        }
        double alpha = 1.0;
        if (mode == VBOC_MODE_SQP && o.globalization) alpha = line_search(ls.it, st.ls_evals);
#pragma unroll 1
        for (int k = 0; k <= N; ++k) {
            for (int i = 0; i < NZ; ++i) s(k, L::Z + i) += alpha * s(k, L::DZ + i);
            if (k < N)
                for (int i = 0; i < NX; ++i) s(k, L::PI + i) = (1.0 - alpha) * s(k, L::PI + i) + alpha * s(k, L::PIQ + i);
            for (int c = 0; c < NC; ++c) s(k, L::LAM + c) = (1.0 - alpha) * s(k, L::LAM + c) + alpha * s(k, L::LAMQ + c);
        }
        ++ls.it;
        return false;
    }

    VB_HD void begin(LaneState &ls, const Prob &pb) {
        load_problem(pb);
        ls.have = 1, ls.it = 0;
        ls.st.status = VBOC_MAXITER, ls.st.sqp_iter = 0, ls.st.qp_iter = 0, ls.st.ls_evals = 0, ls.st.qp_status = 0;
        ls.st.pad_ = 0, ls.st.cost = 0.0;
        ls.st.res_stat = ls.st.res_eq = ls.st.res_ineq = ls.st.res_comp = 0.0;
    }
    VB_HD void finish(LaneState &ls, const Prob &pb) {
        ls.st.cost = total_cost(L::Z);
        store_solution(pb, ls.st);
        ls.have = 0;
    }
};

}  // namespace vboc
