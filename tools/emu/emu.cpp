// emu.cpp -- host emulation of the warp solver (development / test harness, NOT the product).
// Compiles vboc_b200/csrc/ocp_warp.h with VBOC_EMU so that the 32 lanes of every lane region run
// as a loop; lets the lane program be debugged against the oracle on a machine without a GPU.
// Nothing under vboc_b200/ loads this library.
#define VBOC_EMU
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../vboc_b200/csrc/ocp_warp.h"
#include "../../vboc_b200/csrc/ocp_lane.h"
#include "../../vboc_b200/csrc/datagen_warp.h"

using namespace vboc;

// optional multiplier export of the warp solver (set before emu_solve_batch, reset to null afterwards)
static double *g_pi = nullptr, *g_lam = nullptr;
extern "C" void emu_set_multiplier_out(double *pi, double *lam) { g_pi = pi, g_lam = lam; }
// optional in-kernel guess network of the AL family (W2T / W3T already transposed by the caller) + guess export
static GuessNet g_gnn;
static bool g_gnn_on = false;
static double *g_xg_out = nullptr;
extern "C" void emu_set_guess_net(int on, int hidden, int n_out, const double *W1, const double *b1, const double *W2T,
                                  const double *b2, const double *W3T, const double *b3, double mean, double stdv,
                                  double *xg_out) {
    g_gnn_on = on != 0, g_xg_out = xg_out;
    g_gnn.hidden = hidden, g_gnn.n_out = n_out, g_gnn.W1 = W1, g_gnn.b1 = b1, g_gnn.W2T = W2T, g_gnn.b2 = b2;
    g_gnn.W3T = W3T, g_gnn.b3 = b3, g_gnn.mean = mean, g_gnn.stdv = stdv;
}

// optional Cartesian path constraint of the VBOC family (n = 2): WarpSolver<2, VBOC_FAMILY_CART>, row export
static bool g_cart_on = false;
static double g_cart[4] = {0, 0, 0, 0};  // xc, yc, lh, uh
static double *g_rowm = nullptr;
extern "C" void emu_set_cartesian(int on, double xc, double yc, double lh, double uh, double *rowm) {
    g_cart_on = on != 0, g_cart[0] = xc, g_cart[1] = yc, g_cart[2] = lh, g_cart[3] = uh, g_rowm = rowm;
}

template <int NQ, int FAM>
static void run(int mode, int batch, int Nmax, const int *N, const double *xg, const double *ug,
                const double *p, const double *lbx0, const double *ubx0, const double *lbx,
                const double *ubx, const double *lbxN, const double *ubxN, const double *lbu,
                const double *ubu, const double *dir, const double *h, const vboc_opts *o, double *x,
                double *u, vboc_stats *st) {
    constexpr bool VB = FAM == VBOC_FAMILY_VBOC || FAM == VBOC_FAMILY_CART;
    const int nxr = 2 * NQ + VB, nu = NQ;
#pragma omp parallel
    {
        std::vector<double> buf(Work<NQ>::doubles_rows(Nmax));
        Smem<NQ> *sm = new Smem<NQ>();
        SmemMpc<NQ> *gm = new SmemMpc<NQ>();
#pragma omp for schedule(dynamic, 1)
        for (int b = 0; b < batch; ++b) {
            Work<NQ> w;
            w.carve(buf.data(), Nmax);
            Prob pb;
            pb.N = N[b], pb.nxr = nxr, pb.h = h[b];
            pb.wt = p ? p[(size_t)b * (NQ + 1) + NQ] : 0.0;
            pb.xg = xg + (size_t)b * (Nmax + 1) * nxr, pb.ug = ug + (size_t)b * Nmax * nu;
            pb.p = p ? p + (size_t)b * (NQ + 1) : nullptr;
            pb.lbx0 = lbx0 + (size_t)b * nxr, pb.ubx0 = ubx0 + (size_t)b * nxr;
            pb.lbx = lbx + (size_t)b * nxr, pb.ubx = ubx + (size_t)b * nxr;
            pb.lbxN = lbxN + (size_t)b * nxr, pb.ubxN = ubxN + (size_t)b * nxr;
            pb.lbu = lbu + (size_t)b * nu, pb.ubu = ubu + (size_t)b * nu;
            pb.dir = dir ? dir + (size_t)b * NQ : nullptr;
            pb.x = x + (size_t)b * (Nmax + 1) * nxr, pb.u = u + (size_t)b * Nmax * nu;
            pb.st = st + b;
            if (g_pi) pb.pi_out = g_pi + (size_t)b * Nmax * 2 * NQ, pb.lam_out = g_lam + (size_t)b * (Nmax + 1) * 6 * NQ;
            if (FAM == VBOC_FAMILY_AL && g_gnn_on) {
                pb.gnn = &g_gnn;
                pb.xg_out = g_xg_out ? g_xg_out + (size_t)b * (Nmax + 1) * nxr : nullptr;
            }
            if constexpr (FAM == VBOC_FAMILY_CART) {
                pb.cart_xc = g_cart[0], pb.cart_yc = g_cart[1], pb.lh = g_cart[2], pb.uh = g_cart[3];
                pb.rowm_out = g_rowm ? g_rowm + (size_t)b * (Nmax + 1) * 6 : nullptr;
            }
            WarpSolver<NQ, FAM> sol(*sm, w, *o, gm);
            sol.solve(pb, mode);
        }
        delete sm;
        delete gm;
    }
}

// the lane-per-OCP solver with W = 1 (one "lane" at a time)
template <int NQ, int FAM, int DTS = 0>
static void run_lane(int mode, int batch, int Nmax, const int *N, const double *xg, const double *ug,
                     const double *p, const double *lbx0, const double *ubx0, const double *lbx,
                     const double *ubx, const double *lbxN, const double *ubxN, const double *lbu,
                     const double *ubu, const double *dir, const double *h, const vboc_opts *o, double *x,
                     double *u, vboc_stats *st) {
    const int nxr = 2 * NQ + (FAM == VBOC_FAMILY_VBOC), nu = NQ;
#pragma omp parallel
    {
        std::vector<double> buf(LaneLayout<NQ, DTS>::TOTAL);
#pragma omp for schedule(dynamic, 1)
        for (int b = 0; b < batch; ++b) {
            Prob pb;
            pb.N = N[b], pb.nxr = nxr, pb.h = h[b];
            pb.wt = p ? p[(size_t)b * (NQ + 1) + NQ] : 0.0;
            pb.xg = xg + (size_t)b * (Nmax + 1) * nxr, pb.ug = ug + (size_t)b * Nmax * nu;
            pb.p = p ? p + (size_t)b * (NQ + 1) : nullptr;
            pb.lbx0 = lbx0 + (size_t)b * nxr, pb.ubx0 = ubx0 + (size_t)b * nxr;
            pb.lbx = lbx + (size_t)b * nxr, pb.ubx = ubx + (size_t)b * nxr;
            pb.lbxN = lbxN + (size_t)b * nxr, pb.ubxN = ubxN + (size_t)b * nxr;
            pb.lbu = lbu + (size_t)b * nu, pb.ubu = ubu + (size_t)b * nu;
            pb.dir = dir ? dir + (size_t)b * NQ : nullptr;
            pb.x = x + (size_t)b * (Nmax + 1) * nxr, pb.u = u + (size_t)b * Nmax * nu;
            pb.st = st + b;
            double scratch[LaneSolver<NQ, FAM, 1, DTS>::SM_TOTAL];
            LaneSolver<NQ, FAM, 1, DTS> sol(buf.data(), scratch, 0, *o);
            LaneState ls;
            sol.begin(ls, pb);
            while (!sol.sqp_iteration(ls, mode)) {
            }
            sol.finish(ls, pb);
        }
    }
}

extern "C" int emu_lane_solve_batch(int n, int family, int mode, int batch, int Nmax, const int *N,
                                    const double *xg, const double *ug, const double *p,
                                    const double *lbx0, const double *ubx0, const double *lbx,
                                    const double *ubx, const double *lbxN, const double *ubxN,
                                    const double *lbu, const double *ubu, const double *dir,
                                    const double *h, const vboc_opts *o, double *x, double *u,
                                    vboc_stats *st) {
#define GO(NQ, FAM)                                                                                   \
    if (n == NQ && family == FAM) {                                                                   \
        run_lane<NQ, FAM>(mode, batch, Nmax, N, xg, ug, p, lbx0, ubx0, lbx, ubx, lbxN, ubxN, lbu, ubu, \
                          dir, h, o, x, u, st);                                                       \
        return 0;                                                                                     \
    }
    GO(1, 0) GO(2, 0) GO(3, 0) GO(1, 1) GO(2, 1) GO(3, 1)
#undef GO
    return -1;
}

// 1-DOF VBOC with the dt state kept (free dt)
extern "C" int emu_lane_dts_solve_batch(int n, int family, int mode, int batch, int Nmax, const int *N,
                                        const double *xg, const double *ug, const double *p,
                                        const double *lbx0, const double *ubx0, const double *lbx,
                                        const double *ubx, const double *lbxN, const double *ubxN,
                                        const double *lbu, const double *ubu, const double *dir,
                                        const double *h, const vboc_opts *o, double *x, double *u,
                                        vboc_stats *st) {
    if (n != 1 || family != 0) return -1;
    run_lane<1, 0, 1>(mode, batch, Nmax, N, xg, ug, p, lbx0, ubx0, lbx, ubx, lbxN, ubxN, lbu, ubu, dir, h, o, x, u,
                      st);
    return 0;
}

extern "C" int emu_solve_batch(int n, int family, int mode, int batch, int Nmax, const int *N,
                               const double *xg, const double *ug, const double *p,
                               const double *lbx0, const double *ubx0, const double *lbx,
                               const double *ubx, const double *lbxN, const double *ubxN,
                               const double *lbu, const double *ubu, const double *dir,
                               const double *h, const vboc_opts *o, double *x, double *u,
                               vboc_stats *st) {
#define GO(NQ, FAM)                                                                              \
    if (n == NQ && family == FAM) {                                                              \
        run<NQ, FAM>(mode, batch, Nmax, N, xg, ug, p, lbx0, ubx0, lbx, ubx, lbxN, ubxN, lbu, ubu, \
                     dir, h, o, x, u, st);                                                       \
        return 0;                                                                                \
    }
    if (g_cart_on) {
        if (n != 2 || family != VBOC_FAMILY_VBOC) return -1;
        run<2, VBOC_FAMILY_CART>(mode, batch, Nmax, N, xg, ug, p, lbx0, ubx0, lbx, ubx, lbxN, ubxN, lbu, ubu, dir, h, o, x, u, st);
        return 0;
    }
    GO(1, 0) GO(2, 0) GO(3, 0) GO(1, 1) GO(2, 1) GO(3, 1)
#undef GO
    return -1;
}

// the device-resident data generation state machine (datagen_warp.h) on the host
template <int NQ>
static void run_datagen(int count, const DgParams &P, const vboc_opts *o, const int *js, const double *p, const double *lb0,
                        const double *ub0, const double *retry, double *rows, DgCounters *cnt) {
#pragma omp parallel
    {
        std::vector<double> buf(Work<NQ>::TOTAL + DgWork<NQ>::TOTAL);
        Smem<NQ> *sm = new Smem<NQ>();
#pragma omp for schedule(dynamic, 1)
        for (int b = 0; b < count; ++b) {
            Work<NQ> w;
            w.carve(buf.data(), DG_N_CAP);
            DgWork<NQ> g;
            g.carve(buf.data() + Work<NQ>::TOTAL);
            WarpSolver<NQ, VBOC_FAMILY_VBOC> sol(*sm, w, *o);
            DataGen<NQ> dg(sol, g, P);
            DgIO<NQ> io{js, p, lb0, ub0, retry, rows, cnt};
            dg.run(io, b);
            cnt[b] = dg.c;
        }
        delete sm;
    }
}

extern "C" int emu_datagen_run(int n, int count, int N0, double dt, double tol, const int *js, const double *p,
                               const double *lb0, const double *ub0, const double *retry, const vboc_opts *o,
                               double *rows, DgCounters *cnt) {
    DgParams P;
    P.N0 = N0, P.dt = dt, P.tol = tol;
    P.q_min = M_PI - M_PI / 4, P.q_max = M_PI + M_PI / 4, P.v_max = 10.0, P.u_max = 10.0;
    if (n == 2) run_datagen<2>(count, P, o, js, p, lb0, ub0, retry, rows, cnt);
    else if (n == 3) run_datagen<3>(count, P, o, js, p, lb0, ub0, retry, rows, cnt);
    else return -1;
    return 0;
}

// MPC family: tracking cost + the learned margin as a terminal constraint
template <int NQ>
static void run_mpc(int mode, int batch, int Nmax, const int *N, const double *xg, const double *ug, const double *x0,
                    const double *lbx, const double *ubx, const double *lbu, const double *ubu, const double *Wz,
                    const double *WzN, const double *yref, const double *yrefN, double Tf, const NnNet &net, double lh,
                    double uh, const vboc_opts *o, double *x, double *u, vboc_stats *st, double *lamg, int rows_soft,
                    const double *rowZ, double *rowm) {
    const int nxr = 2 * NQ, nu = NQ, nz = 3 * NQ;
#pragma omp parallel
    {
        std::vector<double> buf(Work<NQ>::doubles_rows(Nmax));
        Smem<NQ> *sm = new Smem<NQ>();
        SmemMpc<NQ> *gm = new SmemMpc<NQ>();
#pragma omp for schedule(dynamic, 1)
        for (int b = 0; b < batch; ++b) {
            Work<NQ> w;
            w.carve(buf.data(), Nmax);
            Prob pb;
            pb.N = N[b], pb.nxr = nxr, pb.h = Tf / N[b], pb.wt = 0.0, pb.p = nullptr;
            pb.xg = xg + (size_t)b * (Nmax + 1) * nxr, pb.ug = ug + (size_t)b * Nmax * nu;
            pb.lbx0 = x0 + (size_t)b * nxr, pb.ubx0 = x0 + (size_t)b * nxr;
            pb.lbx = lbx, pb.ubx = ubx, pb.lbxN = lbx, pb.ubxN = ubx, pb.lbu = lbu, pb.ubu = ubu, pb.dir = nullptr;
            pb.x = x + (size_t)b * (Nmax + 1) * nxr, pb.u = u + (size_t)b * Nmax * nu, pb.st = st + b;
            pb.Wz = Wz, pb.WzN = WzN, pb.yref = yref + (size_t)b * nz, pb.yrefN = yrefN + (size_t)b * nxr;
            pb.nn = &net, pb.lh = lh, pb.uh = uh, pb.lamg_out = lamg ? lamg + 2 * b : nullptr;
            pb.rows_soft = rows_soft, pb.rowZ = rowZ ? rowZ + (size_t)b * (Nmax + 1) * 4 : nullptr;
            pb.rowm_out = rowm ? rowm + (size_t)b * (Nmax + 1) * 6 : nullptr;
            if (g_pi) pb.pi_out = g_pi + (size_t)b * Nmax * 2 * NQ, pb.lam_out = g_lam + (size_t)b * (Nmax + 1) * 6 * NQ;
            WarpSolver<NQ, VBOC_FAMILY_MPC> sol(*sm, w, *o, gm);
            sol.solve(pb, mode);
        }
        delete sm;
        delete gm;
    }
}

extern "C" int emu_solve_mpc(int n, int mode, int batch, int Nmax, const int *N, const double *xg, const double *ug,
                             const double *x0, const double *lbx, const double *ubx, const double *lbu, const double *ubu,
                             const double *Wz, const double *WzN, const double *yref, const double *yrefN, double Tf,
                             int hidden, const double *W1, const double *b1, const double *W2, const double *b2,
                             const double *W3, double b3, double mean, double stdv, double scale, double lh, double uh,
                             const vboc_opts *o, double *x, double *u, vboc_stats *st, double *lamg, int rows_soft,
                             const double *rowZ, double *rowm, int vstart) {
    std::vector<double> w2t((size_t)hidden * hidden);
    for (int j = 0; j < hidden; ++j)
        for (int k = 0; k < hidden; ++k) w2t[(size_t)k * hidden + j] = W2[(size_t)j * hidden + k];
    NnNet net;
    net.n_in = 2 * n, net.hidden = hidden, net.W1 = W1, net.b1 = b1, net.W2 = W2, net.W2T = w2t.data(), net.b2 = b2;
    net.W3 = W3, net.b3 = b3, net.mean = mean, net.stdv = stdv, net.scale = scale, net.vstart = vstart < 0 ? n : vstart;
    if (hidden > NN_HMAX) return -1;
    if (n == 2) run_mpc<2>(mode, batch, Nmax, N, xg, ug, x0, lbx, ubx, lbu, ubu, Wz, WzN, yref, yrefN, Tf, net, lh, uh, o, x, u, st, lamg, rows_soft, rowZ, rowm);
    else if (n == 3) run_mpc<3>(mode, batch, Nmax, N, xg, ug, x0, lbx, ubx, lbu, ubu, Wz, WzN, yref, yrefN, Tf, net, lh, uh, o, x, u, st, lamg, rows_soft, rowZ, rowm);
    else return -1;
    return 0;
}

template <int NQ>
static void run_testdata(int count, const DgParams &P, const vboc_opts *o, const TestIO<NQ> &io) {
#pragma omp parallel
    {
        std::vector<double> buf(Work<NQ>::TOTAL + DgWork<NQ>::TOTAL);
        Smem<NQ> *sm = new Smem<NQ>();
#pragma omp for schedule(dynamic, 1)
        for (int b = 0; b < count; ++b) {
            Work<NQ> w;
            w.carve(buf.data(), DG_N_CAP);
            DgWork<NQ> g;
            g.carve(buf.data() + Work<NQ>::TOTAL);
            WarpSolver<NQ, VBOC_FAMILY_VBOC> sol(*sm, w, *o);
            DataGen<NQ> dg(sol, g, P);
            dg.run_testing(io, b);
            io.cnt[b] = dg.c;
        }
        delete sm;
    }
}

extern "C" int emu_testdata_run(int n, int count, int N0, double dt, int max_solves, const double *ran, const double *q_init,
                                const double *retry, const vboc_opts *o, double *rows, DgCounters *cnt) {
    DgParams P;
    P.N0 = N0, P.dt = dt, P.tol = 1e-3;
    P.q_min = M_PI - M_PI / 4, P.q_max = M_PI + M_PI / 4, P.v_max = 10.0, P.u_max = 10.0;
    if (n == 2) {
        TestIO<2> io{ran, q_init, retry, rows, cnt, max_solves, 4, 1e-4};
        run_testdata<2>(count, P, o, io);
    } else if (n == 3) {
        TestIO<3> io{ran, q_init, retry, rows, cnt, max_solves, 3, 1e-3};
        run_testdata<3>(count, P, o, io);
    } else return -1;
    return 0;
}
