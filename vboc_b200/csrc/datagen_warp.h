// datagen_warp.h -- the per-problem state machine of the reference's VBOC data generation ON THE DEVICE.
//
// Stands in for `data_generation(v)` (VBOC/triplependulum_vboc.py:19-370, VBOC/doublependulum_vboc.py:19-402): one
// warp carries one problem from its sampled data to its list of saved rows without going back to the host --
//   * the extreme trajectory: at most 10 solves, horizon + 1 while the cost still decreases by more than tol,
//     after a failed solve the direction / free positions are perturbed and the solve restarted (:107-174);
//   * the walk along it (:205-365): every state is classified as on the boundary of the viability kernel, on a state
//     limit, or inside, with sub-OCPs from the state (chains of <= 5 solves with growing horizon, :232-289) and the
//     simulated "unviable twin" (one RK4 step of the reference's AcadosSimSolver per state, :347-352);
//   * the row filter (:362-365).
// Every solve is WarpSolver::solve (ocp_warp.h), the same code the batched C-ABI entry points run.
//
// The reference draws its random numbers from the unseeded `random` module inside the worker.  Here the host
// draws them from the per-problem Philox stream (vboc_b200/drivers.py: the same stream, in the same order, as the
// host generators `drivers.data_generation_worker`) and hands them over as data: the sampled problem and the
// perturbations of up to 10 restarts.  Given the same seed the device returns the rows of the host generators.
#pragma once
#include "ocp_warp.h"

namespace vboc {

constexpr int DG_N_CAP = 128;                 // largest horizon (drivers.N_CAP)
constexpr int DG_ROWS_MAX = 2 * DG_N_CAP + 2; // rows one problem can save
constexpr int DG_RETRIES = 10;

// per-problem input (host prepared), output and counters
struct DgParams {
    int N0;
    double dt, tol, q_min, q_max, v_max, u_max;
};
struct DgCounters {
    int status;    // 0 rows returned, 1 no extreme trajectory (the generator returns None), 2 row buffer overflow
    int n_rows, solves, converged, sim_steps, sqp_iter, qp_iter;
    int t_done_us;  // when the problem finished, microseconds after the kernel started (device clock)
};

template <int NQ>
struct DgIO {
    static constexpr int NXR = 2 * NQ + 1;
    // inputs, one row per problem
    const int *joint_sel;   // [count]
    const double *p;        // [count][NQ + 1]
    const double *lb0, *ub0;  // [count][NXR]
    const double *retry;    // [count][DG_RETRIES][NQ + 1]: perturbation of the direction (NQ) and of the free positions
    // outputs
    double *rows;           // [count][DG_ROWS_MAX][2 NQ]
    DgCounters *cnt;        // [count]
};

// inputs / outputs of the test-data state machine (`testing(v)`, triplependulum_testdata.py:9-125)
constexpr int TG_MAX_SOLVES = 60;
template <int NQ>
struct TestIO {
    const double *ran;     // [count][NQ]                   un-normalised cost direction
    const double *q_init;  // [count][NQ]                   initial position
    const double *retry;   // [count][TG_MAX_SOLVES][2 NQ]  per restart: perturbation of ran, then of q_init
    double *rows;          // [count][2 NQ]                 the boundary state x_0 (status 0)
    DgCounters *cnt;       // [count]
    int max_solves, digits;
    double cost_tol;
};

// scratch of one warp slot in global memory (reference-shaped solver inputs / outputs and the walk's trajectory)
template <int NQ>
struct DgWork {
    static constexpr int NXR = 2 * NQ + 1, NX = 2 * NQ, NU = NQ;
    double *xg, *ug, *xs, *us, *xsol, *usol;
    double *p, *lbx0, *ubx0, *lbx, *ubx, *lbxN, *ubxN, *lbu, *ubu, *dir;
    vboc_stats *st;
    static constexpr size_t S1 = (size_t)(DG_N_CAP + 2) * NXR, S2 = (size_t)(DG_N_CAP + 2) * NU;
    static constexpr size_t TOTAL = ((2 * S1 + 2 * S2 + (size_t)(DG_N_CAP + 2) * NX + S2 + 8 * NXR + 4 * NU + 16 + 16) + 1) & ~(size_t)1;
    VB_HD void carve(double *b) {
        xg = b, ug = xg + S1, xs = ug + S2, us = xs + S1, xsol = us + S2, usol = xsol + (size_t)(DG_N_CAP + 2) * NX;
        p = usol + S2, lbx0 = p + NXR, ubx0 = lbx0 + NXR, lbx = ubx0 + NXR, ubx = lbx + NXR, lbxN = ubx + NXR,
        ubxN = lbxN + NXR, lbu = ubxN + NXR, ubu = lbu + NU, dir = ubu + NU;
        st = reinterpret_cast<vboc_stats *>(dir + NU + (NU & 1) + 2);
    }
};

template <int NQ>
struct DataGen {
    static constexpr int NX = 2 * NQ, NU = NQ, NXR = 2 * NQ + 1;
    WarpSolver<NQ, VBOC_FAMILY_VBOC> &sol;
    DgWork<NQ> g;
    const DgParams &P;
    DgCounters c;

    VB_DEV DataGen(WarpSolver<NQ, VBOC_FAMILY_VBOC> &s_, const DgWork<NQ> &g_, const DgParams &P_) : sol(s_), g(g_), P(P_) {}

    // ---- small helpers (warp-uniform reads; writes inside lane regions)
    VB_DEV double nudge(double q) const {
        const double eps = 10.0 * P.tol;
        q = q > P.q_max - eps ? q - eps : q;
        return q < P.q_min + eps ? q + eps : q;
    }
    VB_DEV double vnorm(const double *v) const {
        double a = 0.0;
#pragma unroll
        for (int i = 0; i < NQ; ++i) a += v[i] * v[i];
        return sqrt(a);
    }
    VB_DEV bool v_out_of_box(const double *x) const {
        bool o = false;
#pragma unroll
        for (int i = 0; i < NQ; ++i) o = o || x[NQ + i] > P.v_max || x[NQ + i] < -P.v_max;
        return o;
    }
    VB_DEV bool q_near_limit(const double *x) const {
        const double eps = 10.0 * P.tol;
        bool o = false;
#pragma unroll
        for (int i = 0; i < NQ; ++i) o = o || x[i] > P.q_max - eps || x[i] < P.q_min + eps;
        return o;
    }
    VB_DEV void load_state(const double *src, double *x) const {
#pragma unroll
        for (int i = 0; i < NX; ++i) x[i] = src[i];
    }

    // constant problem data: bounds of the intermediate / terminal stages and of the controls
    VB_DEV void set_constants() {
        FOR_LANES
        if (lane < NXR) {
            const bool q = lane < NQ, v = lane >= NQ && lane < NX;
            g.lbx[lane] = q ? P.q_min : (v ? -P.v_max : P.dt);
            g.ubx[lane] = q ? P.q_max : (v ? P.v_max : P.dt);
            g.lbxN[lane] = q ? P.q_min : (v ? 0.0 : P.dt);
            g.ubxN[lane] = q ? P.q_max : (v ? 0.0 : P.dt);
        }
        if (lane < NU) g.lbu[lane] = -P.u_max, g.ubu[lane] = P.u_max;
        END_LANES
    }
    // the cost direction and its unit vector for the stage-0 constraint v_0 || d
    VB_DEV void set_direction(const double *pv) {
        const double nrm = vnorm(pv);
        FOR_LANES
        if (lane < NQ) g.p[lane] = pv[lane], g.dir[lane] = pv[lane] / nrm;
        if (lane == NQ) g.p[NQ] = 0.0;
        END_LANES
    }
    // guess of `ramp_guess(N)` (VBOC/triplependulum_vboc.py:86-93, gravity compensation doublependulum_vboc.py:84),
    // expanded as OCP_solve does: stages 0..N-1 from the N rows, stage N repeats the last row
    VB_DEV void ramp_guess(int N, int joint_sel, double q_init_sel, double q_fin_sel) {
        FOR_LANES
        for (int k = lane; k <= N; k += 32) {
            const int kk = k < N ? k : N - 1;
            const double tau = N > 1 ? (kk == N - 1 ? 1.0 : kk * (1.0 / (N - 1))) : 0.0;
            double *x = g.xg + (size_t)k * NXR;
#pragma unroll
            for (int i = 0; i < NQ; ++i) x[i] = g.lbx0[i], x[NQ + i] = 0.0;
            x[NX] = P.dt;
#pragma unroll
            for (int i = 0; i < NQ; ++i)
                if (i == joint_sel) {
                    x[i] = (1.0 - tau) * q_init_sel + tau * q_fin_sel;
                    x[NQ + i] = 2.0 * (1.0 - tau) * (q_fin_sel - q_init_sel);
                }
            if (k < N) {
                double *u = g.ug + (size_t)k * NU;
#pragma unroll
                for (int i = 0; i < NU; ++i)
                    u[i] = NQ == 2 ? PendN::g * PendN::l * (PendN::m * (NQ - i)) * sin(x[i]) : 0.0;
            }
        }
        END_LANES
    }
    // `_extended_guess`: warm start for one more interval from the solution of N_old intervals in (xs, us)
    VB_DEV void extended_guess(int N_old) {
        FOR_LANES
        for (int idx = lane; idx < (N_old + 2) * NXR; idx += 32) {
            const int k = idx / NXR, i = idx - k * NXR;
            g.xg[idx] = g.xs[(size_t)(k <= N_old ? k : N_old) * NXR + i];
        }
        for (int idx = lane; idx < (N_old + 1) * NU; idx += 32) g.ug[idx] = idx < N_old * NU ? g.us[idx] : 0.0;
        END_LANES
    }

    // not inlined: the extreme-trajectory loop and the sub-OCP loop share ONE copy of the solver's code
    VB_DEV_NOINLINE int solve(int N, const double *lb0, const double *ub0) {
        Prob pb;
        pb.N = N, pb.nxr = NXR, pb.h = P.dt, pb.wt = 0.0;
        pb.xg = g.xg, pb.ug = g.ug, pb.p = g.p;
        pb.lbx0 = lb0, pb.ubx0 = ub0, pb.lbx = g.lbx, pb.ubx = g.ubx, pb.lbxN = g.lbxN, pb.ubxN = g.ubxN;
        pb.lbu = g.lbu, pb.ubu = g.ubu, pb.dir = g.dir;
        pb.x = g.xs, pb.u = g.us, pb.st = g.st;
        sol.solve(pb, VBOC_MODE_SQP);
        UNIFORM_SYNC();
        ++c.solves, c.sqp_iter += g.st->sqp_iter, c.qp_iter += g.st->qp_iter;
        if (g.st->status == 0) ++c.converged;
        return g.st->status;
    }

    VB_DEV void save_row(double *rows, const double *x) {
        if (c.n_rows >= DG_ROWS_MAX) {
            c.status = 2;
            return;
        }
        double *dst = rows + (size_t)c.n_rows * NX;
        FOR_LANES
        if (lane < NX) dst[lane] = x[lane];
        END_LANES
        ++c.n_rows;
    }

    // the whole of testing(v) for problem b (triplependulum_testdata.py:9-125, doublependulum_testdata.py:10-121):
    // maximise the initial velocity along a random direction from a random position; horizon + 1 while the rounded
    // cost still decreases by more than cost_tol; after a failed solve restart with direction and position perturbed.
    // Deliberate deviation shared with the host generator: at most max_solves solves instead of `while True`.
    VB_DEV void run_testing(const TestIO<NQ> &io, int b) {
        c.status = 1, c.n_rows = 0, c.solves = 0, c.converged = 0, c.sim_steps = 0, c.sqp_iter = 0, c.qp_iter = 0, c.t_done_us = 0;
        const double *retry = io.retry + (size_t)b * TG_MAX_SOLVES * 2 * NQ;
        set_constants();
        double ran[NQ], qi[NQ];
#pragma unroll
        for (int i = 0; i < NQ; ++i) ran[i] = io.ran[(size_t)b * NQ + i], qi[i] = io.q_init[(size_t)b * NQ + i];
        double scale = 1.0;
        for (int d = 0; d < io.digits; ++d) scale *= 10.0;
        int N = P.N0, n_retry = 0;
        double cost = 1e6;
        bool fresh = true;
#pragma unroll 1
        for (int attempt = 0; attempt < io.max_solves; ++attempt) {
            {
                const double nr = vnorm(ran);
                double d[NQ];
#pragma unroll
                for (int i = 0; i < NQ; ++i) d[i] = ran[i] / nr;
                set_direction(d);
            }
            FOR_LANES
            if (lane < NXR) {
                const bool q = lane < NQ, v = lane >= NQ && lane < NX;
                double qv = 0.0;
#pragma unroll
                for (int i = 0; i < NQ; ++i)
                    if (i == lane) qv = qi[i];
                g.lbx0[lane] = q ? qv : (v ? -P.v_max : P.dt);
                g.ubx0[lane] = q ? qv : (v ? P.v_max : P.dt);
            }
            END_LANES
            if (fresh) {
                // constant guess at the initial position, gravity-compensation torques for the double pendulum
                FOR_LANES
                for (int k = lane; k <= N; k += 32) {
                    double *x = g.xg + (size_t)k * NXR;
#pragma unroll
                    for (int i = 0; i < NQ; ++i) x[i] = qi[i], x[NQ + i] = 0.0;
                    x[NX] = P.dt;
                    if (k < N) {
                        double *u = g.ug + (size_t)k * NU;
#pragma unroll
                        for (int i = 0; i < NU; ++i)
                            u[i] = NQ == 2 ? PendN::g * PendN::l * (PendN::m * (NQ - i)) * sin(qi[i]) : 0.0;
                    }
                }
                END_LANES
                fresh = false;
            }
            const int status = solve(N, g.lbx0, g.ubx0);
            if (status == 0) {
                const double cnew = g.st->cost;
                // python round(cost, digits): nearest multiple of 10^-digits
                if (cnew > nearbyint(cost * scale) / scale - io.cost_tol) {
                    FOR_LANES
                    if (lane < NX) io.rows[(size_t)b * NX + lane] = g.xs[lane];
                    END_LANES
                    c.status = 0, c.n_rows = 1;
                    return;
                }
                if (N + 1 > DG_N_CAP) return;
                cost = cnew;
                extended_guess(N);
                N += 1;
            } else {
                N = P.N0;
                const double *rp = retry + (size_t)n_retry * 2 * NQ;
                ++n_retry;
#pragma unroll
                for (int i = 0; i < NQ; ++i) ran[i] += rp[i], qi[i] += rp[NQ + i];
                fresh = true;
                cost = 1e6;
            }
        }
    }

    // the whole of data_generation(v) for problem b
    VB_DEV void run(const DgIO<NQ> &io, int b) {
        const double eps = 10.0 * P.tol, tol = P.tol;
        c.status = 1, c.n_rows = 0, c.solves = 0, c.converged = 0, c.sim_steps = 0, c.sqp_iter = 0, c.qp_iter = 0, c.t_done_us = 0;
        const int joint_sel = io.joint_sel[b];
        double *rows = io.rows + (size_t)b * DG_ROWS_MAX * NX;
        const double *retry = io.retry + (size_t)b * DG_RETRIES * (NQ + 1);
        set_constants();
        FOR_LANES
        if (lane < NXR) g.lbx0[lane] = io.lb0[(size_t)b * NXR + lane], g.ubx0[lane] = io.ub0[(size_t)b * NXR + lane];
        END_LANES
        double pv[NQ];
#pragma unroll
        for (int i = 0; i < NQ; ++i) pv[i] = io.p[(size_t)b * (NQ + 1) + i];
        set_direction(pv);
        // the selected joint starts AT a position limit (+- eps) and sweeps to the other one
        const bool from_min = g.lbx0[joint_sel] < 0.5 * (P.q_min + P.q_max);
        const double q_init_sel = from_min ? P.q_min : P.q_max, q_fin_sel = from_min ? P.q_max : P.q_min;

        // ---- extreme trajectory (VBOC/triplependulum_vboc.py:107-174)
        int N = P.N0, n_retry = 0;
        double cost = 1e6;
        bool have = false;
        ramp_guess(N, joint_sel, q_init_sel, q_fin_sel);
#pragma unroll 1
        for (int attempt = 0; attempt < 10; ++attempt) {
            const int status = solve(N, g.lbx0, g.ubx0);
            if (status == 0) {
                const double cnew = g.st->cost;
                if (cnew > cost - tol) {
                    have = true;
                    break;
                }
                if (N + 1 > DG_N_CAP) break;
                cost = cnew;
                extended_guess(N);
                N += 1;
            } else {
                N = P.N0;
                const double *rp = retry + (size_t)n_retry * (NQ + 1);
                ++n_retry;
                double d[NQ];
#pragma unroll
                for (int i = 0; i < NQ; ++i) d[i] = g.p[i] + rp[i];
                const double nd = vnorm(d);
#pragma unroll
                for (int i = 0; i < NQ; ++i) d[i] = d[i] / nd;
                set_direction(d);
                const double dev = rp[NQ];
                FOR_LANES
                if (lane < NQ && lane != joint_sel) {
                    const double q = nudge(g.lbx0[lane] + dev);
                    g.lbx0[lane] = q, g.ubx0[lane] = q;
                }
                END_LANES
                ramp_guess(N, joint_sel, q_init_sel, q_fin_sel);
                cost = 1e6;
            }
        }
        if (!have) return;  // status 1: the generator returns None
        c.status = 0;

        // ---- the walk (VBOC/triplependulum_vboc.py:205-365)
        FOR_LANES
        for (int idx = lane; idx < (N + 1) * NX; idx += 32) {
            const int k = idx / NX, i = idx - k * NX;
            g.xsol[idx] = g.xs[(size_t)k * NXR + i];
        }
        for (int idx = lane; idx < N * NU; idx += 32) g.usol[idx] = g.us[idx];
        END_LANES
        save_row(rows, g.xsol);
        double pdir[NQ], xsym[NX], xo[NX];
#pragma unroll
        for (int i = 0; i < NQ; ++i) pdir[i] = g.p[i];
        load_state(g.xsol, xo);
#pragma unroll
        for (int i = 0; i < NQ; ++i) xo[NQ + i] -= eps * pdir[i];
        bool at_limit = v_out_of_box(xo);
        if (!at_limit) {
#pragma unroll
            for (int i = 0; i < NX; ++i) xsym[i] = xo[i];
        }
#pragma unroll 1
        for (int f = 1; f < N; ++f) {
            const double *xf = g.xsol + (size_t)f * NX;
            if (at_limit) {
                load_state(xf, xo);
                {
                    const double nv = vnorm(xo + NQ);
#pragma unroll
                    for (int i = 0; i < NQ; ++i) xo[NQ + i] += eps * xo[NQ + i] / nv;
                }
                if (q_near_limit(xf) || v_out_of_box(xo)) {
                    at_limit = true;
                } else {
                    at_limit = false;
                    if (q_near_limit(xf - NX)) break;  // leaving a position limit the trajectory usually enters the kernel
                    // sub-OCP from x_sol[f]: is the rest of the trajectory on the boundary or inside?
                    int N_t = N - f;
                    double vf[NQ];
#pragma unroll
                    for (int i = 0; i < NQ; ++i) vf[i] = xf[NQ + i];
                    const double norm_old = vnorm(vf);
#pragma unroll
                    for (int i = 0; i < NQ; ++i) pdir[i] = -vf[i] / norm_old;
                    set_direction(pdir);
#pragma unroll
                    for (int i = 0; i < NQ; ++i) pdir[i] = g.p[i];
                    // bounds of the sub-OCP's first stage live behind the solver outputs' tail (scratch rows)
                    double *lbs = g.xg + (size_t)(DG_N_CAP + 1) * NXR, *ubs = g.xs + (size_t)(DG_N_CAP + 1) * NXR;
                    FOR_LANES
                    if (lane < NXR) {
                        const bool q = lane < NQ, v = lane >= NQ && lane < NX;
                        lbs[lane] = q ? xf[lane] : (v ? -P.v_max : P.dt);
                        ubs[lane] = q ? xf[lane] : (v ? P.v_max : P.dt);
                    }
                    for (int idx = lane; idx < (N_t + 1) * NXR; idx += 32) {
                        const int k = idx / NXR, i = idx - k * NXR;
                        g.xg[idx] = i < NX ? g.xsol[(size_t)(f + k) * NX + i] : P.dt;
                    }
                    for (int idx = lane; idx < N_t * NU; idx += 32) g.ug[idx] = g.usol[(size_t)f * NU + idx];
                    END_LANES
                    double norm_bef = 0.0, norm_new = 0.0;
                    bool ok = false;
                    int N_sub = 0;  // horizon of the last converged sub-solve (its solution stays in xs / us)
#pragma unroll 1
                    for (int t = 0; t < 5; ++t) {
                        if (solve(N_t, lbs, ubs) != 0) break;
                        N_sub = N_t;
                        norm_new = vnorm(g.xs + NQ);
                        if (norm_new < norm_bef + tol) {
                            ok = true;
                            break;
                        }
                        if (N_t + 1 > DG_N_CAP) break;
                        norm_bef = norm_new;
                        extended_guess(N_t);
                        N_t += 1;
                    }
                    (void)N_sub;
                    if (ok) {
                        if (norm_new > norm_old + tol) {  // the state is inside the kernel: adopt the better tail
                            FOR_LANES
                            for (int idx = lane; idx < (N - f) * NX; idx += 32) {
                                const int k = idx / NX, i = idx - k * NX;
                                g.xsol[(size_t)(f + k) * NX + i] = g.xs[(size_t)k * NXR + i];
                            }
                            for (int idx = lane; idx < (N - f) * NU; idx += 32) g.usol[(size_t)f * NU + idx] = g.us[idx];
                            END_LANES
                            load_state(xf, xo);
#pragma unroll
                            for (int i = 0; i < NQ; ++i) xo[NQ + i] += eps * xo[NQ + i] / norm_new;
                            at_limit = v_out_of_box(xo);
                            if (!at_limit) {
#pragma unroll
                                for (int i = 0; i < NX; ++i) xsym[i] = xo[i];
                            }
                        } else {  // on the boundary: the unviable twin lies in the cost direction
                            load_state(xf, xo);
#pragma unroll
                            for (int i = 0; i < NQ; ++i) xo[NQ + i] -= eps * pdir[i];
#pragma unroll
                            for (int i = 0; i < NQ; ++i)
                                if (i == joint_sel) xo[NQ + i] = fmin(fmax(xo[NQ + i], -P.v_max), P.v_max);
#pragma unroll
                            for (int i = 0; i < NX; ++i) xsym[i] = xo[i];
                        }
                    } else {  // undecided: keep the state once per later state at a velocity limit, then stop
#pragma unroll 1
                        for (int r = f; r < N; ++r) {
                            const double *xr = g.xsol + (size_t)r * NX;
                            bool lim = false;
#pragma unroll
                            for (int i = 0; i < NQ; ++i) lim = lim || fabs(xr[NQ + i]) > P.v_max - eps;
                            if (lim) save_row(rows, xf);
                        }
                        break;
                    }
                }
            } else {
                // the unviable twin follows the optimal controls: one RK4 step of the simulator (:347-352)
                double un[NU], xn[NX];
#pragma unroll
                for (int i = 0; i < NU; ++i) un[i] = g.usol[(size_t)(f - 1) * NU + i];
                rk4_step<NQ, double>(xsym, un, P.dt, xn);
                ++c.sim_steps;
                bool lim = v_out_of_box(xn);
#pragma unroll
                for (int i = 0; i < NQ; ++i) lim = lim || xn[i] > P.q_max || xn[i] < P.q_min;
#pragma unroll
                for (int i = 0; i < NX; ++i) xsym[i] = xn[i];
                at_limit = lim;
            }
            bool keep = !q_near_limit(xf);
#pragma unroll
            for (int i = 0; i < NQ; ++i) keep = keep && fabs(xf[NQ + i]) > tol;
            if (keep) save_row(rows, xf);
        }
    }
};

}  // namespace vboc
