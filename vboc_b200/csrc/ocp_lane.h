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

#ifndef VB_LANE_PREFETCH
#define VB_LANE_PREFETCH 0
#endif

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

    static constexpr int SMS = W == 1 ? 1 : 128;  // stride of the per-thread shared-memory scratch (= CTA size)
    double *base;
    double *sm;  // per-thread scratch: element idx at sm[idx * SMS]
    int lane;
    const vboc_opts &o;
    // per-OCP scalars
    int N, fixed0, fixedN, termfix, nact;
    int nb, bidx[NX];  // border mode: number and indices of the fixed terminal components
    double h, wtdt, wt, wcost[NQ];

    VB_HD LaneSolver(double *base_, double *sm_, int lane_, const vboc_opts &o_)
        : base(base_), sm(sm_), lane(lane_), o(o_), N(0) {}

    VB_HD double &g(size_t off) const { return base[off * W + lane]; }
    VB_HD double &s(int k, size_t f) const { return base[((size_t)k * L::SREC + f) * W + lane]; }

    // The 32 lanes' copies of one field element share one 256-byte line: lane l prefetches line l of the
    // range into L2, so a stage costs each lane a handful of prefetch instructions.
    VB_HD void prefetch_l2(int k, size_t f0, size_t f1) const {
        if (W == 1 || !VB_LANE_PREFETCH || k < 0 || k > N) return;
        for (size_t f = f0 + lane; f < f1; f += 32) VB_PREFETCH_L2(base + ((size_t)k * L::SREC + f) * W);
    }

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
            prefetch_l2(k + 2, L::Z, L::Z + NZ);
            prefetch_l2(k + 2, L::BD, L::BA + NX * NZ);
            prefetch_l2(k + 3, L::DZ, L::DV);
            const double *sk = &s(k, 0);
            double *skw = &s(k, 0);
            const int sc = sclass(k);
            double r[NZ], v[NZ], z[NZ];
            // block A: loads of [B A], the multipliers and the iterate, then r = H v + g + [B A]'pi - pi_{k-1}
            // and the dynamics residual
            {
                double c[NX][NZ], pi[NX], pim[NX], dzn[NX], bd[NX];
#pragma unroll
                for (int i = 0; i < NZ; ++i) v[i] = sk[(L::DZ + i) * W], z[i] = sk[(L::Z + i) * W];
#pragma unroll
                for (int i = 0; i < NX; ++i) {
                    pi[i] = k < N ? sk[(L::PIQ + i) * W] : 0.0;
                    pim[i] = k > 0 ? s(k - 1, L::PIQ + i) : 0.0;
                    dzn[i] = k < N ? s(k + 1, L::DZ + NU + i) : 0.0;
                    bd[i] = k < N ? sk[(L::BD + i) * W] : 0.0;
#pragma unroll
                    for (int j = 0; j < NZ; ++j) c[i][j] = k < N ? sk[(L::BA + i * NZ + j) * W] : 0.0;
                }
#pragma unroll
                for (int i = 0; i < NZ; ++i) {
                    double a = cost_h(k, i) * v[i] + cost_g(k, i, z[i]);
#pragma unroll
                    for (int m = 0; m < NX; ++m) a += c[m][i] * pi[m];
                    if (i >= NU) a -= pim[i - NU];
                    r[i] = a;
                }
                if (k < N) {
                    double rb[NX];
#pragma unroll
                    for (int i = 0; i < NX; ++i) {
                        double a = bd[i] - dzn[i];
#pragma unroll
                        for (int j = 0; j < NZ; ++j) a += c[i][j] * v[j];
                        rb[i] = a;
                        bad |= (a != a);
                        vb = fmax(vb, fabs(a));
                    }
#pragma unroll
                    for (int i = 0; i < NX; ++i) skw[(L::RB + i) * W] = rb[i];
                }
            }
            // block B: the bound constraints
            double hh[NZ], bar[NZ];
            {
                double ll[NZ], lu[NZ], tl[NZ], tu[NZ], bl[NZ], bu[NZ];
#pragma unroll
                for (int i = 0; i < NZ; ++i) {
                    ll[i] = sk[(L::LAMQ + i) * W], lu[i] = sk[(L::LAMQ + NZ + i) * W];
                    tl[i] = sk[(L::TQ + i) * W], tu[i] = sk[(L::TQ + NZ + i) * W];
                    bl[i] = g(L::LB + sc * NZ + i), bu[i] = g(L::UB + sc * NZ + i);
                }
#pragma unroll
                for (int i = 0; i < NZ; ++i) {
                    hh[i] = cost_h(k, i) + o.qp_reg_prim, bar[i] = 0.0;
                    if (active(k, i)) {
                        r[i] += lu[i] - ll[i];
                        double dl = (bl[i] - z[i]) - v[i] + tl[i], du = v[i] - (bu[i] - z[i]) + tu[i];
                        double ml = ll[i] * tl[i], mu_ = lu[i] * tu[i];
                        bad |= (dl != dl) | (du != du) | (ml != ml) | (mu_ != mu_);
                        vd = fmax(vd, fmax(fabs(dl), fabs(du)));
                        vm = fmax(vm, fmax(fabs(ml), fabs(mu_)));
                        mu += ml + mu_;
                        double itl = 1.0 / tl[i], itu = 1.0 / tu[i];
                        hh[i] += ll[i] * itl + lu[i] * itu;
                        bar[i] = (ml - ll[i] * dl) * itl - (mu_ - lu[i] * du) * itu;
                    } else if (k == N) {
                        r[i] = 0.0;
                    }
                }
            }
            if (k == 0) proj0(r + NU);
#pragma unroll
            for (int i = 0; i < NZ; ++i) {
                bad |= (r[i] != r[i]);
                vg = fmax(vg, fabs(r[i]));
                skw[(L::HH + i) * W] = hh[i];
                skw[(L::RR + i) * W] = r[i] + bar[i];
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
    // fixed terminal component bidx[bj] (unit terminal gradient, all other right-hand sides zero).
    //
    // Register plan of the factor step (n = 3: 54 + 21 + 6 + 6 + 6 doubles live at the peak): the stage's
    // [B A] block and the symmetric P+ stay in registers, T_j = P+ c_j is formed one column at a time, the
    // entries M[a][j] = c_a . T_j go straight to the per-thread shared-memory scratch `sm` (45 + 9 doubles),
    // from which the nu x nu Cholesky and the Schur complement read them back.  No per-thread arrays in
    // local memory on this path.
    VB_HD static constexpr int TRI(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }
    static constexpr int SM_M = 0, SM_V = NZ * (NZ + 1) / 2, SM_B = SM_V + NZ, SM_TOTAL = SM_B + NZ;
    VB_HD double &scr(int idx) const { return sm[idx * SMS]; }

    VB_HD bool backward(int mode, double sm_, double *dx0, int bj = 0) {
        const bool factor = mode == 0;
        double P[NX * (NX + 1) / 2], pv[NX];
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            bool fx = termfix && ((fixedN >> i) & 1);
            double hh = s(N, L::HH + NU + i), rr = rhs_of(N, NU + i, mode, sm_);
            if (mode == 3) rr = (i == bidx[bj]) ? 1.0 : 0.0;
            g(L::HHN + i) = fx ? 0.0 : hh;
            g(L::RN + i) = fx ? 0.0 : rr;
            pv[i] = fx ? 0.0 : rr;
#pragma unroll
            for (int j = 0; j <= i; ++j) P[TRI(i, j)] = (j == i && !fx) ? hh : 0.0;
        }
        bool ok = true;
#pragma unroll 1
        for (int k = N - 1; k >= 0; --k) {
            const bool last = (k == N - 1) && termfix;
            prefetch_l2(k - 2, L::BA, factor ? L::Q1 : L::P);
            const double *sk = &s(k, 0);  // stage base: every field is sk[field * W]
            double *skw = &s(k, 0);
            double c[NX][NZ];
#pragma unroll
            for (int i = 0; i < NX; ++i)
#pragma unroll
                for (int j = 0; j < NZ; ++j) c[i][j] = sk[(L::BA + i * NZ + j) * W];
            // right-hand sides of the stage, loaded before any store of this iteration
            double rhsv[NZ];
#pragma unroll
            for (int j = 0; j < NZ; ++j) rhsv[j] = rhs_of(k, j, mode, sm_);
            if (factor) {
                double tb[NX], hhv[NZ], rb[NX];
#pragma unroll
                for (int j = 0; j < NZ; ++j) hhv[j] = sk[(L::HH + j) * W];
#pragma unroll
                for (int i = 0; i < NX; ++i) rb[i] = sk[(L::RB + i) * W];
#pragma unroll
                for (int i = 0; i < NX; ++i) {
                    double a = 0.0;
#pragma unroll
                    for (int mm = 0; mm < NX; ++mm) a += P[TRI(i, mm)] * rb[mm];
                    tb[i] = a;
                }
#pragma unroll
                for (int j = 0; j < NZ; ++j) {
                    double T[NX];
#pragma unroll
                    for (int i = 0; i < NX; ++i) {
                        double t = 0.0;
#pragma unroll
                        for (int mm = 0; mm < NX; ++mm) t += P[TRI(i, mm)] * c[mm][j];
                        T[i] = t;
                    }
#pragma unroll
                    for (int a_ = 0; a_ <= j; ++a_) {
                        double t = (a_ == j) ? hhv[j] : 0.0;
#pragma unroll
                        for (int mm = 0; mm < NX; ++mm) t += c[mm][a_] * T[mm];
                        scr(SM_M + TRI(j, a_)) = t;
                    }
                    double mb = 0.0, mp = 0.0;
#pragma unroll
                    for (int i = 0; i < NX; ++i) mb += c[i][j] * tb[i], mp += c[i][j] * pv[i];
                    scr(SM_B + j) = mb;
                    scr(SM_V + j) = mb + mp + rhsv[j];
                }
#pragma unroll
                for (int j = 0; j < NZ; ++j) skw[(L::MB + j) * W] = scr(SM_B + j);
            } else {
                double mbv[NZ];
#pragma unroll
                for (int j = 0; j < NZ; ++j) mbv[j] = mode == 3 ? 0.0 : sk[(L::MB + j) * W];
#pragma unroll
                for (int j = 0; j < NZ; ++j) {
                    double t = mbv[j] + rhsv[j];
#pragma unroll
                    for (int i = 0; i < NX; ++i) t += c[i][j] * pv[i];
                    scr(SM_V + j) = t;
                }
                if (last)
#pragma unroll
                    for (int a_ = 0; a_ < NZ; ++a_)
#pragma unroll
                        for (int b_ = 0; b_ <= a_; ++b_) scr(SM_M + TRI(a_, b_)) = g(L::MF + a_ * NZ + b_);
            }
            if (!last) {
                double Lu[NU][NU], di[NU], y[NU];
                if (factor) {
#pragma unroll
                    for (int j = 0; j < NU; ++j) {
                        double d = scr(SM_M + TRI(j, j));
#pragma unroll
                        for (int cc = 0; cc < j; ++cc) d -= Lu[j][cc] * Lu[j][cc];
                        di[j] = d > 0.0 ? VB_RSQRT(d) : 0.0;
#pragma unroll
                        for (int i = j + 1; i < NU; ++i) {
                            double a = scr(SM_M + TRI(i, j));
#pragma unroll
                            for (int cc = 0; cc < j; ++cc) a -= Lu[i][cc] * Lu[j][cc];
                            Lu[i][j] = a * di[j];
                        }
                    }
#pragma unroll
                    for (int i = 0; i < NU; ++i)
#pragma unroll
                        for (int cc = 0; cc <= i; ++cc) skw[(L::LUU + i * NU + cc) * W] = (cc == i) ? di[i] : Lu[i][cc];
                } else {
#pragma unroll
                    for (int i = 0; i < NU; ++i) {
                        di[i] = sk[(L::LUU + i * NU + i) * W];
#pragma unroll
                        for (int cc = 0; cc < i; ++cc) Lu[i][cc] = sk[(L::LUU + i * NU + cc) * W];
                    }
                }
#pragma unroll
                for (int cc = 0; cc < NU; ++cc) {
                    double a = scr(SM_V + cc);
#pragma unroll
                    for (int c2 = 0; c2 < cc; ++c2) a -= Lu[cc][c2] * y[c2];
                    y[cc] = a * di[cc];
                    skw[(L::YV + cc) * W] = y[cc];
                }
                double Lxu[NX][NU];
#pragma unroll
                for (int i = 0; i < NX; ++i) {
#pragma unroll
                    for (int cc = 0; cc < NU; ++cc) {
                        double a;
                        if (factor) {
                            a = scr(SM_M + TRI(NU + i, cc));
#pragma unroll
                            for (int c2 = 0; c2 < cc; ++c2) a -= Lxu[i][c2] * Lu[cc][c2];
                            a *= di[cc];
                            skw[(L::LXU + i * NU + cc) * W] = a;
                        } else {
                            a = sk[(L::LXU + i * NU + cc) * W];
                        }
                        Lxu[i][cc] = a;
                    }
                    double pn = scr(SM_V + NU + i);
#pragma unroll
                    for (int cc = 0; cc < NU; ++cc) pn -= Lxu[i][cc] * y[cc];
                    pv[i] = pn;
                    skw[(L::PV + i) * W] = pn;
                }
                if (factor) {
#pragma unroll
                    for (int i = 0; i < NX; ++i)
#pragma unroll
                        for (int j = 0; j <= i; ++j) {
                            double a = scr(SM_M + TRI(NU + i, NU + j));
#pragma unroll
                            for (int cc = 0; cc < NU; ++cc) a -= Lxu[i][cc] * Lxu[j][cc];
                            P[TRI(i, j)] = a;
                        }
#pragma unroll
                    for (int i = 0; i < NX; ++i)
#pragma unroll
                        for (int j = 0; j < NX; ++j) skw[(L::P + i * NX + j) * W] = P[TRI(i, j)];
                }
            } else {
                // terminal velocity equalities through the last control: du = K dx + k0
                double Gi[NU][NU], K[NU][NX], k0[NU], tmp[NZ];
                if (factor) {
                    double G[NU][NU];
                    for (int a = 0; a < NU; ++a)
                        for (int b2 = 0; b2 < NU; ++b2) G[a][b2] = c[NQ + a][b2];
                    ok = inverse_small(G, Gi) && ok;
                    for (int a = 0; a < NU; ++a)
                        for (int j = 0; j < NX; ++j) {
                            double v = 0.0;
                            for (int b2 = 0; b2 < NU; ++b2) v -= Gi[a][b2] * c[NQ + b2][NU + j];
                            K[a][j] = v;
                            s(k, L::LXU + j * NU + a) = v;
                        }
                    for (int a = 0; a < NU; ++a)
                        for (int b2 = 0; b2 < NU; ++b2) s(k, L::LUU + a * NU + b2) = Gi[a][b2];
                    for (int a_ = 0; a_ < NZ; ++a_)
                        for (int b_ = 0; b_ < NZ; ++b_) g(L::MF + a_ * NZ + b_) = scr(SM_M + TRI(a_, b_));
                } else {
                    for (int a = 0; a < NU; ++a) {
                        for (int b2 = 0; b2 < NU; ++b2) Gi[a][b2] = s(k, L::LUU + a * NU + b2);
                        for (int j = 0; j < NX; ++j) K[a][j] = s(k, L::LXU + j * NU + a);
                    }
                }
                for (int a = 0; a < NU; ++a) {
                    double v = 0.0;
                    for (int b2 = 0; b2 < NU; ++b2) v -= Gi[a][b2] * (g(L::EN + NQ + b2) + s(k, L::RB + NQ + b2));
                    k0[a] = v;
                    g(L::K0 + a) = v;
                }
                for (int i = 0; i < NZ; ++i) {
                    double v = scr(SM_V + i);
                    for (int a = 0; a < NU; ++a) v += scr(SM_M + TRI(i, a)) * k0[a];
                    tmp[i] = v;
                    g(L::ML + i) = scr(SM_V + i);
                }
                if (factor) {
                    for (int i = 0; i < NX; ++i)
                        for (int j = 0; j <= i; ++j) {
                            double a = scr(SM_M + TRI(NU + i, NU + j));
                            for (int cc = 0; cc < NU; ++cc) {
                                a += K[cc][i] * scr(SM_M + TRI(cc, NU + j)) + scr(SM_M + TRI(NU + i, cc)) * K[cc][j];
                                for (int c2 = 0; c2 < NU; ++c2) a += K[cc][i] * scr(SM_M + TRI(cc, c2)) * K[c2][j];
                            }
                            P[TRI(i, j)] = a;
                        }
                    for (int i = 0; i < NX; ++i)
                        for (int j = 0; j < NX; ++j) s(k, L::P + i * NX + j) = P[TRI(i, j)];
                }
                for (int j = 0; j < NX; ++j) {
                    double pn = tmp[NU + j];
                    for (int a = 0; a < NU; ++a) pn += K[a][j] * tmp[a];
                    pv[j] = pn;
                    s(k, L::PV + j) = pn;
                }
            }
        }
        // stage 0: dx0 = -e0 + Z0 dy, (Z0'P0 Z0) dy = -Z0'(p0 - P0 e0)
        if (factor) {
            double T[NX][NX], Lz[NX][NX], dzi[NX];
            for (int i = 0; i < NX; ++i) {
                double a = 0.0;
                for (int j = 0; j < NX; ++j) a += P[TRI(i, j)] * g(L::E0 + j);
                g(L::PE + i) = a;
                for (int cc = 0; cc < NX; ++cc) {
                    double t = 0.0;
                    for (int j = 0; j < NX; ++j) t += P[TRI(i, j)] * g(L::Z0 + j * NX + cc);
                    T[i][cc] = t;
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
                for (int cc = 0; cc < j; ++cc) d -= Lz[j][cc] * Lz[j][cc];
                dzi[j] = d > 0.0 ? VB_RSQRT(d) : 0.0;
                for (int i = j + 1; i < NX; ++i) {
                    double a = Lz[i][j];
                    for (int cc = 0; cc < j; ++cc) a -= Lz[i][cc] * Lz[j][cc];
                    Lz[i][j] = a * dzi[j];
                }
            }
            for (int i = 0; i < NX; ++i) {
                g(L::DZI + i) = dzi[i];
                for (int cc = 0; cc < i; ++cc) g(L::LZ + i * NX + cc) = Lz[i][cc];
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
                for (int cc = 0; cc < NX; ++cc) a += g(L::Z0 + i * NX + cc) * dy[cc];
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
    // Memory discipline of the streaming loops: every access goes through `base`, so the compiler cannot move a
    // load above an earlier store.  Each block therefore LOADS everything it needs first (independent loads, one
    // exposed round trip), computes in registers, and stores last.
    static constexpr int CBS = (NZ % 3 == 0) ? 3 : 2;  // components per block (code size vs loads in flight)
    VB_HD void con_block(int k, int I0, const double *dzb, int mode, double sm, double &al, double &s0, double &s1,
                         double &s2) {
        constexpr int NI = CBS;
        const double *sk = &s(k, 0);
        double *skw = &s(k, 0);
        const int sc = sclass(k);
        double z[NI], v[NI], lam[2][NI], t[2][NI], pr[2][NI], rr[NI], bl[NI], bu[NI];
#pragma unroll
        for (int a = 0; a < NI; ++a) {
            const int i = I0 + a;
            z[a] = sk[(L::Z + i) * W], v[a] = sk[(L::DZ + i) * W];
            bl[a] = g(L::LB + sc * NZ + i), bu[a] = g(L::UB + sc * NZ + i);
            rr[a] = mode == 0 ? sk[(L::RR + i) * W] : 0.0;
#pragma unroll
            for (int sd = 0; sd < 2; ++sd) {
                lam[sd][a] = sk[(L::LAMQ + sd * NZ + i) * W], t[sd][a] = sk[(L::TQ + sd * NZ + i) * W];
                pr[sd][a] = mode == 1 ? sk[(L::PROD + sd * NZ + i) * W] : 0.0;
            }
        }
        double o1[2][NI], o2[2][NI], q1[NI], q2[NI];
#pragma unroll
        for (int a = 0; a < NI; ++a) {
            const int i = I0 + a;
            q1[a] = 0.0, q2[a] = 0.0;
            const bool act = active(k, i);
#pragma unroll
            for (int sd = 0; sd < 2; ++sd) {
                const double lm = act ? lam[sd][a] : 1.0, tt = act ? t[sd][a] : 1.0;
                const double rd = sd ? v[a] - (bu[a] - z[a]) + tt : (bl[a] - z[a]) - v[a] + tt;
                double rm = lm * tt;
                if (mode == 1) rm += pr[sd][a] - sm;
                if (mode == 2) rm -= sm;
                const double dtt = (sd ? -dzb[a] : dzb[a]) - rd;
                const double it = 1.0 / tt;
                const double dl = -(rm + lm * dtt) * it;
                if (act) {
                    if (dtt < 0.0 && tt + al * dtt < 0.0) al = fmin(al, -tt / dtt);
                    if (dl < 0.0 && lm + al * dl < 0.0) al = fmin(al, -lm / dl);
                    s0 += lm * tt, s1 += lm * dtt + tt * dl, s2 += dtt * dl;
                    const double p_ = dtt * dl;
                    q1[a] += sd ? -p_ * it : p_ * it;
                    q2[a] += sd ? -it : it;
                    o1[sd][a] = mode == 0 ? p_ : dtt, o2[sd][a] = dl;
                } else {
                    o1[sd][a] = 0.0, o2[sd][a] = 0.0;
                }
            }
        }
#pragma unroll
        for (int a = 0; a < NI; ++a) {
            const int i = I0 + a;
            if (active(k, i)) {
#pragma unroll
                for (int sd = 0; sd < 2; ++sd) {
                    if (mode == 0) {
                        skw[(L::PROD + sd * NZ + i) * W] = o1[sd][a];
                    } else {
                        skw[(L::DT + sd * NZ + i) * W] = o1[sd][a], skw[(L::DLAM + sd * NZ + i) * W] = o2[sd][a];
                    }
                }
            }
            if (mode == 0) skw[(L::Q1 + i) * W] = rr[a] + q1[a], skw[(L::Q2 + i) * W] = q2[a];
        }
    }
    // slack / multiplier steps of the bound constraints of stage k for the primal step dz: ratio test and
    // the sums of mu(alpha); mode 0 stores the second-order products and Q1, Q2, modes 1/2 store DT, DLAM
    VB_HD void con_stage(int k, const double *dz, int mode, double sm, double &al, double &s0, double &s1,
                         double &s2) {
#pragma unroll 1
        for (int b0 = 0; b0 < NZ / CBS; ++b0) {
            double dzb[CBS];
#pragma unroll
            for (int a = 0; a < CBS; ++a) {
                // dz lives in registers: pick the block's entries with selects instead of dynamic indexing
                double v = 0.0;
#pragma unroll
                for (int i = 0; i < NZ; ++i)
                    if (i == b0 * CBS + a) v = dz[i];
                dzb[a] = v;
            }
            con_block(k, b0 * CBS, dzb, mode, sm, al, s0, s1, s2);
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
#pragma unroll
        for (int i = 0; i < NX; ++i) dx[i] = dx0[i];
#pragma unroll 1
        for (int k = 0; k <= N; ++k) {
            const bool last = (k == N - 1) && termfix;
            prefetch_l2(k + 2, L::Z, L::Z + NZ);
            prefetch_l2(k + 2, L::BA, L::BA + NX * NZ + NX);
            prefetch_l2(k + 2, L::LUU, L::DV);
            if (mode == 0) prefetch_l2(k + 2, L::RR, L::Q1);
            if (mode == 1) prefetch_l2(k + 2, L::PROD, L::DLAM);
            const double *sk = &s(k, 0);
            double *skw = &s(k, 0);
            double dz[NZ];
#pragma unroll
            for (int i = 0; i < NX; ++i) dz[NU + i] = dx[i];
#pragma unroll
            for (int cc = 0; cc < NU; ++cc) dz[cc] = 0.0;
            if (k < N) {
                // phase 1: the control step from the factors
                double lxu[NX][NU], luu[NU][NU], yv[NU];
#pragma unroll
                for (int j = 0; j < NX; ++j)
#pragma unroll
                    for (int cc = 0; cc < NU; ++cc) lxu[j][cc] = sk[(L::LXU + j * NU + cc) * W];
#pragma unroll
                for (int i = 0; i < NU; ++i) {
                    yv[i] = last ? g(L::K0 + i) : sk[(L::YV + i) * W];
#pragma unroll
                    for (int cc = 0; cc <= i; ++cc) luu[i][cc] = sk[(L::LUU + i * NU + cc) * W];
                }
                if (!last) {
                    double t[NU];
#pragma unroll
                    for (int cc = 0; cc < NU; ++cc) {
                        double a = yv[cc];
#pragma unroll
                        for (int j = 0; j < NX; ++j) a += lxu[j][cc] * dx[j];
                        t[cc] = a;
                    }
#pragma unroll
                    for (int cc = NU - 1; cc >= 0; --cc) {
                        double a = -t[cc];
#pragma unroll
                        for (int c2 = cc + 1; c2 < NU; ++c2) a -= luu[c2][cc] * dz[c2];
                        dz[cc] = a * luu[cc][cc];
                    }
                } else {
#pragma unroll
                    for (int a_ = 0; a_ < NU; ++a_) {
                        double a = yv[a_];
#pragma unroll
                        for (int j = 0; j < NX; ++j) a += lxu[j][a_] * dx[j];
                        dz[a_] = a;
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < NZ; ++i) skw[(fDV + i) * W] = dz[i];
            if (!raw) con_stage(k, dz, mode, sm, al, s0, s1, s2);
            if (k < N) {
                // phase 2: the next state
                double dxn[NX];
                {
                    double c[NX][NZ], rb[NX];
#pragma unroll
                    for (int i = 0; i < NX; ++i) {
                        rb[i] = mode == 3 ? 0.0 : sk[(L::RB + i) * W];
#pragma unroll
                        for (int j = 0; j < NZ; ++j) c[i][j] = sk[(L::BA + i * NZ + j) * W];
                    }
#pragma unroll
                    for (int i = 0; i < NX; ++i) {
                        double a = rb[i];
#pragma unroll
                        for (int j = 0; j < NZ; ++j) a += c[i][j] * dz[j];
                        dxn[i] = a;
                    }
                }
                // phase 3: dpi_k = P_{k+1} dx_{k+1} + p_{k+1}
                if (k + 1 < N) {
                    const double *sn = &s(k + 1, 0);
                    double Pn[NX][NX], pn[NX], dpi[NX];
#pragma unroll
                    for (int i = 0; i < NX; ++i) {
                        pn[i] = sn[(L::PV + i) * W];
#pragma unroll
                        for (int j = 0; j < NX; ++j) Pn[i][j] = sn[(L::P + i * NX + j) * W];
                    }
#pragma unroll
                    for (int i = 0; i < NX; ++i) {
                        double a = pn[i];
#pragma unroll
                        for (int j = 0; j < NX; ++j) a += Pn[i][j] * dxn[j];
                        dpi[i] = a;
                    }
#pragma unroll
                    for (int i = 0; i < NX; ++i) skw[(fDPI + i) * W] = dpi[i];
                } else {
                    double nuv[NU];
                    if (last) {
                        double tu[NU];
                        for (int a_ = 0; a_ < NU; ++a_) {
                            double a = g(L::ML + a_);
                            for (int j = 0; j < NZ; ++j) a += g(L::MF + a_ * NZ + j) * dz[j];
                            tu[a_] = a;
                        }
                        for (int b2 = 0; b2 < NU; ++b2) {
                            double a = 0.0;
                            for (int a_ = 0; a_ < NU; ++a_) a -= s(k, L::LUU + a_ * NU + b2) * tu[a_];
                            nuv[b2] = a;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < NX; ++i) {
                        double a = g(L::HHN + i) * dxn[i] + g(L::RN + i);
                        if (last && i >= NQ) a = nuv[i - NQ];
                        skw[(fDPI + i) * W] = a;
                    }
                }
#pragma unroll
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
                if (!DTS || nb == 0) {
                    alpha = forward(ph, sm, dx0, S0, S1, S2);
                } else if constexpr (DTS != 0) {
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
                prefetch_l2(k + 2, L::DZ, L::PROD);
                prefetch_l2(k + 2, L::DLAM, L::WDYN);
                const double *sk = &s(k, 0);
                double *skw = &s(k, 0);
                {
                    double a0[NZ], a1[NZ], b0[NX], b1[NX];
#pragma unroll
                    for (int i = 0; i < NZ; ++i) a0[i] = sk[(L::DZ + i) * W], a1[i] = sk[(L::DV + i) * W];
#pragma unroll
                    for (int i = 0; i < NX; ++i) b0[i] = sk[(L::PIQ + i) * W], b1[i] = sk[(L::DPI + i) * W];
#pragma unroll
                    for (int i = 0; i < NZ; ++i) skw[(L::DZ + i) * W] = a0[i] + as * a1[i];
                    if (k < N)
#pragma unroll
                        for (int i = 0; i < NX; ++i) skw[(L::PIQ + i) * W] = b0[i] + as * b1[i];
                }
                {
                    double a0[NC], a1[NC];
#pragma unroll
                    for (int c = 0; c < NC; ++c) a0[c] = sk[(L::LAMQ + c) * W], a1[c] = sk[(L::DLAM + c) * W];
#pragma unroll
                    for (int c = 0; c < NC; ++c)
                        if (active(k, c >= NZ ? c - NZ : c)) skw[(L::LAMQ + c) * W] = fmax(a0[c] + as * a1[c], o.qp_lam_min);
#pragma unroll
                    for (int c = 0; c < NC; ++c) a0[c] = sk[(L::TQ + c) * W], a1[c] = sk[(L::DT + c) * W];
#pragma unroll
                    for (int c = 0; c < NC; ++c)
                        if (active(k, c >= NZ ? c - NZ : c)) skw[(L::TQ + c) * W] = fmax(a0[c] + as * a1[c], o.qp_t_min);
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
