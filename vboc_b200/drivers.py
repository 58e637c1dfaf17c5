"""Batched per-problem worker logic of the reference's drivers.

The reference runs `data_generation(v)` / `testing(v)` once per problem inside `multiprocessing.Pool`
workers; each worker interleaves host decisions (horizon extension, retries with perturbed data,
the walk along the optimal trajectory with sub-OCPs and the simulated "unviable twin") with blocking
solver calls.  Here every problem is a Python generator that YIELDS its next request -- an OCP to solve
(`SolveReq`, the arguments of `OCP_solve`) or one RK4 step to simulate (`SimReq`) -- and is resumed with
the answer.  `run_workers` advances all live generators round by round and serves each round's
requests with ONE batched GPU call per kind (`BatchSolver.solve`, `sim_step`), so the host control flow
stays per-problem and faithful while the solves of a round run side by side on the GPU.

  testing_worker           triplependulum_testdata.py:9-125, doublependulum_testdata.py:10-121
  data_generation_worker   VBOC/triplependulum_vboc.py:19-370 (VBOC/doublependulum_vboc.py:19-402 has the
                           same structure; its gravity-compensation torque guess :84 is applied for n = 2)

The random draws follow the reference's distributions but come from a per-problem Philox stream
(seed, problem id): the reference uses the unseeded `random` module.
Deliberate deviation: `testing` retries forever in the reference (`while True`); here a problem gives up
after `max_solves` solves and is reported as failed (None): only a point that passed the reference's
convergence test `cost_new > round(cost) - tol` is ever returned.  Likewise a horizon that would grow past N_CAP
ends the problem as failed instead of accepting the capped solve.
"""
import numpy as np

from . import problems as pr
from ._lib import MODE_SQP

N_CAP = 128  # largest horizon the engine is created for (reference horizons stay below ~115)


class SolveReq:
    """Arguments of `OCP_solve` (VBOC/triplependulum_class_vboc.py:155) for the current `ocp.N`."""
    __slots__ = ("N", "x_guess", "u_guess", "p", "q_lb", "q_ub", "u_lb", "u_ub", "q_init_lb", "q_init_ub",
                 "q_fin_lb", "q_fin_ub")

    def __init__(self, N, x_guess, u_guess, p, q_lb, q_ub, u_lb, u_ub, q_init_lb, q_init_ub, q_fin_lb, q_fin_ub):
        self.N, self.x_guess, self.u_guess, self.p = N, x_guess, u_guess, p
        self.q_lb, self.q_ub, self.u_lb, self.u_ub = q_lb, q_ub, u_lb, u_ub
        self.q_init_lb, self.q_init_ub, self.q_fin_lb, self.q_fin_ub = q_init_lb, q_init_ub, q_fin_lb, q_fin_ub


class SimReq:
    """`sim.acados_integrator` set x / u / T, solve, get x (VBOC/triplependulum_vboc.py:347-352)."""
    __slots__ = ("x", "u", "T")

    def __init__(self, x, u, T):
        self.x, self.u, self.T = x, u, T


class SolveAns:
    __slots__ = ("status", "cost", "x", "u")

    def __init__(self, status, cost, x, u):
        self.status, self.cost, self.x, self.u = status, cost, x, u


def _rng(seed, pid):
    return np.random.Generator(np.random.Philox(key=int(seed), counter=[0, 0, 0, int(pid)]))


def _pm(rng):
    return -1.0 if rng.random() < 0.5 else 1.0


def _pack(n, reqs, N_max=N_CAP):
    """SolveReq list -> batched problem dict for `BatchSolver.solve` / `StreamSolver.submit` (stage N takes the last
    guess row); guesses are laid out for a solver created with horizon capacity `N_max`."""
    B, nx = len(reqs), 2 * n + 1
    if any(r.N > N_max for r in reqs):
        raise ValueError(f"a request needs horizon {max(r.N for r in reqs)} but the solver was created for N_max = {N_max}")
    bp = dict(n=n, family="vboc", N=np.array([r.N for r in reqs], dtype=np.int32),
              x_guess=np.zeros((B, N_max + 1, nx)), u_guess=np.zeros((B, N_max, n)),
              p=np.stack([r.p for r in reqs]))
    for b, r in enumerate(reqs):
        xg, ug = pr.expand_guess(r.x_guess, r.u_guess, r.N)
        bp["x_guess"][b, :r.N + 1] = xg
        bp["u_guess"][b, :r.N] = ug
    for key, attr in (("lbx0", "q_init_lb"), ("ubx0", "q_init_ub"), ("lbx", "q_lb"), ("ubx", "q_ub"),
                      ("lbxN", "q_fin_lb"), ("ubxN", "q_fin_ub"), ("lbu", "u_lb"), ("ubu", "u_ub")):
        bp[key] = np.stack([getattr(r, attr) for r in reqs])
    bp["C0"] = pr.stage0_projector(bp["p"][:, :n], nx) if n > 1 else None
    return bp


def run_workers(n, workers, solver, sim_step, stats=None):
    """Advance the generators until all have returned; returns their return values in order.

    solver.solve(bp, mode) -> dict(status, cost, x, u) (a `BatchSolver`, or the oracle in the CPU tests);
    sim_step(n, X, U, T) -> X_next."""
    results = [None] * len(workers)
    pending = {}
    for i, w in enumerate(workers):
        try:
            pending[i] = next(w)
        except StopIteration as e:
            results[i] = e.value
    rounds = 0
    while pending:
        rounds += 1
        answers = {}
        sol_ids = [i for i, r in pending.items() if isinstance(r, SolveReq)]
        sim_ids = [i for i, r in pending.items() if isinstance(r, SimReq)]
        if sol_ids:
            out = solver.solve(_pack(n, [pending[i] for i in sol_ids], getattr(solver, "N_max", N_CAP)), MODE_SQP)
            for b, i in enumerate(sol_ids):
                N = pending[i].N
                answers[i] = SolveAns(int(out["status"][b]), float(out["cost"][b]), out["x"][b, :N + 1].copy(),
                                      out["u"][b, :N].copy())
            if stats is not None:
                stats["solves"] = stats.get("solves", 0) + len(sol_ids)
                stats["converged"] = stats.get("converged", 0) + int((out["status"] == 0).sum())
        if sim_ids:
            T = pending[sim_ids[0]].T
            Xn = sim_step(n, np.stack([pending[i].x for i in sim_ids]), np.stack([pending[i].u for i in sim_ids]), T)
            for b, i in enumerate(sim_ids):
                answers[i] = Xn[b]
        for i, ans in answers.items():
            try:
                pending[i] = workers[i].send(ans)
            except StopIteration as e:
                results[i] = e.value
                del pending[i]
    if stats is not None:
        stats["rounds"] = rounds
    return results


def run_workers_stream(n, workers, ssol, stats=None, max_launch=1024, idle_sleep=5e-5):
    """`run_workers` over a `engine.StreamSolver`: an event loop instead of rounds.  A worker's next solve is
    submitted as soon as the worker has it ready and the worker is resumed as soon as ITS solve is back
    (tickets); RK4 requests of all workers that are ready at the same time go out as one batched call.
    Same workers, same answers as `run_workers` -- only the waiting changes."""
    import time
    results = [None] * len(workers)
    ready = []   # (worker id, request) not yet served
    owner = {}   # ticket -> (worker id, N)
    live = 0

    t_start = time.perf_counter()
    t_done = np.zeros(len(workers))
    t_solve = []

    def advance(i, ans, first=False):
        nonlocal live
        try:
            ready.append((i, next(workers[i]) if first else workers[i].send(ans)))
        except StopIteration as e:
            results[i] = e.value
            t_done[i] = time.perf_counter() - t_start
            live -= 1

    live = len(workers)
    for i in range(len(workers)):
        advance(i, None, first=True)
    nsolve = nconv = nsim = loops = 0
    while live:
        loops += 1
        sims = [(i, r) for i, r in ready if isinstance(r, SimReq)]
        sols = [(i, r) for i, r in ready if isinstance(r, SolveReq)]
        ready.clear()
        # solves first: they run while the host serves the RK4 steps
        if sols:
            k = min(len(sols), ssol.free_slots, max_launch)
            if k:
                tk = ssol.submit(_pack(n, [r for _, r in sols[:k]], getattr(ssol, "N_max", N_CAP)), MODE_SQP)
                for t, (i, r) in zip(tk.tolist(), sols[:k]):
                    owner[t] = (i, r.N)
            ready.extend(sols[k:])  # no free slot yet: next turn
        if sims:
            T = sims[0][1].T
            cap = getattr(ssol, "cap", len(sims))  # vboc_stream_sim_step takes at most `capacity` rows per call
            for lo in range(0, len(sims), cap):
                part = sims[lo:lo + cap]
                Xn = ssol.sim_step(np.stack([r.x for _, r in part]), np.stack([r.u for _, r in part]), T)
                for b, (i, _) in enumerate(part):
                    advance(i, Xn[b])
            nsim += len(sims)
        done = ssol.poll()
        for t in done:
            i, N = owner.pop(t)
            out = ssol.fetch(t)
            nsolve += 1
            nconv += out["status"] == 0
            t_solve.append(time.perf_counter() - t_start)
            advance(i, SolveAns(int(out["status"]), float(out["cost"]), out["x"], out["u"]))
        if not done and not sims and not (sols and ssol.free_slots):
            time.sleep(idle_sleep)
    if stats is not None:
        stats["solves"] = stats.get("solves", 0) + nsolve
        stats["converged"] = stats.get("converged", 0) + int(nconv)
        stats["sim_steps"] = nsim
        stats["loops"] = loops
        # when the workers finished: the last percent is a few problems whose solves hit the iteration limit
        if t_solve:  # throughput while the GPU is still full: until 90 % of the solves are back
            t90 = float(np.percentile(t_solve, 90))
            stats["solves_per_s_first_90pct"] = round(0.9 * len(t_solve) / max(t90, 1e-9), 1)
        stats["t_done_p50_p90_p99_max"] = [round(float(v), 2) for v in np.percentile(t_done, [50, 90, 99, 100])]
    return results


# ------------------------------------------------------------------------------------------------
def _limits(n, mdl, dt_sym):
    q_min, q_max, v_max, tau = mdl.thetamin, mdl.thetamax, mdl.dthetamax, mdl.umax
    q_lb = np.array([q_min] * n + [-v_max] * n + [dt_sym])
    q_ub = np.array([q_max] * n + [v_max] * n + [dt_sym])
    q_fin_lb = np.array([q_min] * n + [0.0] * n + [dt_sym])
    q_fin_ub = np.array([q_max] * n + [0.0] * n + [dt_sym])
    return q_lb, q_ub, np.full(n, -tau), np.full(n, tau), q_fin_lb, q_fin_ub


def _extended_guess(ans, n):
    """Warm start for N+1 intervals: the solution plus its last state / a zero control
    (VBOC/triplependulum_vboc.py:121-129)."""
    return np.vstack([ans.x, ans.x[-1:]]), np.vstack([ans.u, np.zeros((2, n))])[:ans.u.shape[0] + 1]


TG_MAX_SOLVES = 60


def _tg_initial(n, rng, mdl):
    """The sampling block of `testing(v)` (triplependulum_testdata.py:15-24): direction and position."""
    ran = np.array([_pm(rng) * rng.random() for _ in range(n)])
    q_init = mdl.thetamin + rng.random(n) * (mdl.thetamax - mdl.thetamin)
    return ran, q_init


def _tg_retry(n, rng):
    """Draws of one restart (triplependulum_testdata.py:100-121): perturbation of the direction, then of the position."""
    dr = np.array([rng.random() * _pm(rng) * 0.01 for _ in range(n)])
    dq = np.array([rng.random() * _pm(rng) * 0.01 for _ in range(n)])
    return dr, dq


def testing_inputs(n, num_prob, seed, first=0):
    """Inputs of the device test-data state machine (`vboc_testdata_run`): what `testing_worker` draws, in its order."""
    mdl = pr.Model(n)
    out = dict(ran=np.zeros((num_prob, n)), q_init=np.zeros((num_prob, n)), retry=np.zeros((num_prob, TG_MAX_SOLVES, 2 * n)))
    for b in range(num_prob):
        rng = _rng(seed, first + b)
        out["ran"][b], out["q_init"][b] = _tg_initial(n, rng, mdl)
        for r in range(TG_MAX_SOLVES):
            dr, dq = _tg_retry(n, rng)
            out["retry"][b, r, :n], out["retry"][b, r, n:] = dr, dq
    return out


def testing_worker(n, rng, N0=100, dt_sym=1e-2, max_solves=60):
    """One test-set point: maximise the initial velocity along a random direction from a uniformly random
    position; extend the horizon while the (3-decimal rounded) cost still decreases by > cost_tol; on a
    solver failure restart with the direction and position perturbed by <= 0.01."""
    mdl = pr.Model(n)
    # triplependulum_testdata.py:82 compares with the cost rounded to 3 decimals minus 1e-3,
    # doublependulum_testdata.py:80 with 4 decimals minus 1e-4
    digits, cost_tol = (4, 1e-4) if n == 2 else (3, 1e-3)
    q_lb, q_ub, u_lb, u_ub, q_fin_lb, q_fin_ub = _limits(n, mdl, dt_sym)
    ran, q_init = _tg_initial(n, rng, mdl)

    def fresh(N):
        xg = np.tile(np.concatenate([q_init, np.zeros(n), [dt_sym]]), (N, 1))
        ug = np.tile(mdl.gravity_comp(q_init) if n == 2 else np.zeros(n), (N, 1))
        return xg, ug

    N, cost = N0, 1e6
    xg, ug = fresh(N)
    for _ in range(max_solves):
        p = np.concatenate([ran / np.linalg.norm(ran), [0.0]])
        lb0 = np.concatenate([q_init, np.full(n, -mdl.dthetamax), [dt_sym]])
        ub0 = np.concatenate([q_init, np.full(n, mdl.dthetamax), [dt_sym]])
        ans = yield SolveReq(N, xg, ug, p, q_lb, q_ub, u_lb, u_ub, lb0, ub0, q_fin_lb, q_fin_ub)
        if ans.status == 0:
            if ans.cost > round(cost, digits) - cost_tol:
                return ans.x[0, :2 * n].copy()
            if N + 1 > N_CAP:
                return None
            cost = ans.cost
            xg, ug = _extended_guess(ans, n)
            N += 1
        else:
            N = N0
            dr, dq = _tg_retry(n, rng)
            ran, q_init = ran + dr, q_init + dq
            xg, ug = fresh(N)
            cost = 1e6
    return None


def _dg_initial(n, rng, mdl, dt_sym, eps):
    """The sampling block of `data_generation` (VBOC/triplependulum_vboc.py:33-83): selected joint and side, cost
    direction, initial positions.  Shared by the host generator and by the inputs of the device state machine, so
    both consume the per-problem stream identically."""
    q_min, q_max, v_max = mdl.thetamin, mdl.thetamax, mdl.dthetamax

    def nudge(q):
        q = q - eps if q > q_max - eps else q
        return q + eps if q < q_min + eps else q

    joint_sel = int(rng.integers(0, n))
    vel_sel = _pm(rng)
    q_init_sel, q_fin_sel = (q_min, q_max) if vel_sel < 0 else (q_max, q_min)
    ran = np.array([vel_sel * rng.random()] + [_pm(rng) * rng.random() for _ in range(n - 1)])
    ran = ran / np.linalg.norm(ran)
    others = [j for j in range(n) if j != joint_sel]
    p = np.zeros(n + 1)
    p[joint_sel] = ran[0]
    p[others] = ran[1:]
    q0 = np.array([nudge(q_min + rng.random() * (q_max - q_min)) for _ in range(n)])
    lb0 = np.concatenate([q0, np.full(n, -v_max), [dt_sym]])
    ub0 = np.concatenate([q0, np.full(n, v_max), [dt_sym]])
    lb0[joint_sel] = ub0[joint_sel] = q_min + eps if vel_sel < 0 else q_max - eps
    return dict(joint_sel=joint_sel, vel_sel=vel_sel, p=p, lb0=lb0, ub0=ub0, q_init_sel=q_init_sel, q_fin_sel=q_fin_sel)


def _dg_retry(n, rng):
    """Draws of one restart after a failed solve (VBOC/triplependulum_vboc.py:138-174): perturbation of the cost
    direction (n components) and of the free initial positions (one value)."""
    dp = np.array([rng.random() * _pm(rng) * 0.01 for _ in range(n)])
    dev = rng.random() * _pm(rng) * 0.01
    return dp, dev


DG_RETRIES = 10  # the extreme-trajectory loop makes at most 10 solves, so at most 10 restarts


def dg_inputs(n, num_prob, seed, first=0, dt_sym=1e-2, tol=1e-3):
    """Inputs of the device state machine (`vboc_datagen_run`) for problems first .. first + num_prob - 1: everything
    `data_generation_worker` would draw from its per-problem stream, drawn in the same order."""
    mdl = pr.Model(n)
    out = dict(joint_sel=np.zeros(num_prob, dtype=np.int32), p=np.zeros((num_prob, n + 1)),
               lb0=np.zeros((num_prob, 2 * n + 1)), ub0=np.zeros((num_prob, 2 * n + 1)),
               retry=np.zeros((num_prob, DG_RETRIES, n + 1)))
    for b in range(num_prob):
        rng = _rng(seed, first + b)
        ini = _dg_initial(n, rng, mdl, dt_sym, 10 * tol)
        out["joint_sel"][b], out["p"][b], out["lb0"][b], out["ub0"][b] = ini["joint_sel"], ini["p"], ini["lb0"], ini["ub0"]
        for r in range(DG_RETRIES):
            dp, dev = _dg_retry(n, rng)
            out["retry"][b, r, :n], out["retry"][b, r, n] = dp, dev
    return out


def data_generation_worker(n, rng, N0=100, dt_sym=1e-2, tol=1e-3, restarts=True):
    """One VBOC problem: an extreme trajectory from a position limit of a randomly selected joint, then the
    walk along it that classifies every state (on the boundary of the viability kernel, on a state limit,
    or inside) with sub-OCPs and the simulated unviable twin.  Returns the list of saved rows [q, v] or None.

    The worker is written for any n; with restarts=False it is the generic `VBOC/vboc.py:23-402` (`system_sel` = n):
    there the restart after a failed solve is commented out (`:162-213`: a failed solve ends the problem), everything
    else is the per-system scripts' logic written as loops over `system_sel`.  Its velocity test of the walk indexes
    rows instead of columns (`x_sol[system_sel:nx-1]`, `:251`, raises on an array truth value); the per-system scripts'
    test is used.  The unseeded draws differ only in their order."""
    mdl = pr.Model(n)
    eps = 10 * tol
    q_min, q_max, v_max = mdl.thetamin, mdl.thetamax, mdl.dthetamax
    q_lb, q_ub, u_lb, u_ub, q_fin_lb, q_fin_ub = _limits(n, mdl, dt_sym)

    def nudge(q):
        q = q - eps if q > q_max - eps else q
        return q + eps if q < q_min + eps else q

    ini = _dg_initial(n, rng, mdl, dt_sym, eps)
    joint_sel, p, lb0, ub0 = ini["joint_sel"], ini["p"], ini["lb0"], ini["ub0"]
    q_init_sel, q_fin_sel = ini["q_init_sel"], ini["q_fin_sel"]
    others = [j for j in range(n) if j != joint_sel]

    def ramp_guess(N):
        tau = np.linspace(0, 1, N)
        xg = np.tile(np.concatenate([lb0[:n], np.zeros(n), [dt_sym]]), (N, 1))
        xg[:, joint_sel] = (1 - tau) * q_init_sel + tau * q_fin_sel
        xg[:, n + joint_sel] = 2 * (1 - tau) * (q_fin_sel - q_init_sel)
        ug = np.stack([mdl.gravity_comp(x[:n]) for x in xg]) if n == 2 else np.zeros((N, n))
        return xg, ug

    # ---- extreme trajectory: at most 10 solves, horizon + 1 while the cost decreases by more than tol
    N, cost, sol = N0, 1e6, None
    xg, ug = ramp_guess(N)
    for _ in range(10):
        ans = yield SolveReq(N, xg, ug, p, q_lb, q_ub, u_lb, u_ub, lb0, ub0, q_fin_lb, q_fin_ub)
        if ans.status == 0:
            if ans.cost > cost - tol:
                sol = ans
                break
            if N + 1 > N_CAP:
                return None
            cost = ans.cost
            xg, ug = _extended_guess(ans, n)
            N += 1
        elif not restarts:
            return None
        else:
            N = N0
            dp, dev = _dg_retry(n, rng)
            d = p[:n] + dp
            p = np.concatenate([d / np.linalg.norm(d), [0.0]])
            for j in others:
                lb0[j] = ub0[j] = nudge(lb0[j] + dev)
            xg, ug = ramp_guess(N)
            cost = 1e6
    if sol is None:
        return None

    x_sol, u_sol = sol.x[:, :2 * n].copy(), sol.u.copy()
    rows = [x_sol[0].copy()]
    x_sym = [None] * (N + 1)

    def v_out_of_box(x):
        return bool(np.any(x[n:] > v_max) or np.any(x[n:] < -v_max))

    def q_near_limit(x):
        return bool(np.any(x[:n] > q_max - eps) or np.any(x[:n] < q_min + eps))

    x_out = x_sol[0].copy()
    x_out[n:] -= eps * p[:n]
    at_limit = v_out_of_box(x_out)
    if not at_limit:
        x_sym[0] = x_out

    for f in range(1, N):
        if at_limit:
            x_out = x_sol[f].copy()
            x_out[n:] += eps * x_out[n:] / np.linalg.norm(x_out[n:])
            if q_near_limit(x_sol[f]) or v_out_of_box(x_out):
                at_limit = True
            else:
                at_limit = False
                if q_near_limit(x_sol[f - 1]):
                    break  # leaving a position limit the trajectory usually enters the kernel
                # sub-OCP from x_sol[f]: is the rest of the trajectory on the boundary or inside?
                N_t = N - f
                vf = x_sol[f, n:]
                p = np.concatenate([-vf / np.linalg.norm(vf), [0.0]])
                lb_s = np.concatenate([x_sol[f, :n], np.full(n, -v_max), [dt_sym]])
                ub_s = np.concatenate([x_sol[f, :n], np.full(n, v_max), [dt_sym]])
                dtc = np.full((N - f + 1, 1), dt_sym)
                xg = np.hstack([x_sol[f:N + 1], dtc])
                ug = np.vstack([u_sol[f:N], np.zeros((1, n))])
                norm_old, norm_bef, norm_new, ok, sub = np.linalg.norm(vf), 0.0, 0.0, False, None
                for _ in range(5):
                    ans = yield SolveReq(N_t, xg, ug, p, q_lb, q_ub, u_lb, u_ub, lb_s, ub_s, q_fin_lb, q_fin_ub)
                    if ans.status != 0:
                        break
                    sub = ans
                    norm_new = np.linalg.norm(ans.x[0, n:2 * n])
                    if norm_new < norm_bef + tol:
                        ok = True
                        break
                    if N_t + 1 > N_CAP:
                        break
                    norm_bef = norm_new
                    xg, ug = _extended_guess(ans, n)
                    N_t += 1
                if ok:
                    if norm_new > norm_old + tol:  # the state is inside the kernel: adopt the better tail
                        x_sol[f:N] = sub.x[:N - f, :2 * n]
                        u_sol[f:N] = sub.u[:N - f]
                        x_out = x_sol[f].copy()
                        x_out[n:] += eps * x_out[n:] / norm_new
                        at_limit = v_out_of_box(x_out)
                        if not at_limit:
                            x_sym[f] = x_out
                    else:  # on the boundary: the unviable twin lies in the cost direction
                        x_out = x_sol[f].copy()
                        x_out[n:] -= eps * p[:n]
                        x_out[n + joint_sel] = min(max(x_out[n + joint_sel], -v_max), v_max)
                        x_sym[f] = x_out
                else:  # undecided: keep the state once per later state at a velocity limit, then stop
                    for r in range(f, N):
                        if np.any(np.abs(x_sol[r, n:]) > v_max - eps):
                            rows.append(x_sol[f].copy())
                    break
        else:
            x_out = yield SimReq(x_sym[f - 1].copy(), u_sol[f - 1].copy(), dt_sym)
            x_sym[f] = x_out
            at_limit = bool(np.any(x_out[:n] > q_max) or np.any(x_out[:n] < q_min) or v_out_of_box(x_out))
        if not q_near_limit(x_sol[f]) and np.all(np.abs(x_sol[f, n:]) > tol):
            rows.append(x_sol[f].copy())
    return rows


def pendulum_worker(v_sel, N0=50, eps=1e-3):
    """1-DOF VBOC data generation for one side (VBOC/pendulum_vboc.py:54-223): the extreme trajectory with
    a free dt state (horizon + 1 while the initial velocity norm still grows by > 1e-4), then the
    simplified walk: while the state sits on a velocity limit either keep it or re-solve from it.
    Mirrors the driver, except that the status of the sub-OCP solve is the one of that solve (the reference
    tests the stale status of the first solve, :181-183).  Returns rows [q, v]; raises like the driver when
    a solve fails."""
    mdl = pr.Model(1)
    q_min, q_max, v_max, dt = mdl.thetamin, mdl.thetamax, mdl.dthetamax, 1e-2
    if v_sel < 0:
        q_init, q_fin, lb, ub, cost_dir = q_max, q_min, np.array([q_min, -v_max, 0.]), np.array([q_max, 0., dt]), 1.
    else:
        q_init, q_fin, lb, ub, cost_dir = q_min, q_max, np.array([q_min, 0., 0.]), np.array([q_max, v_max, dt]), -1.
    u_lb, u_ub = np.array([-mdl.umax]), np.array([mdl.umax])

    def req(N, xg, ug, q0):
        return SolveReq(N, xg, ug, np.array([cost_dir, 1.]), lb, ub, u_lb, u_ub, np.array([q0, -v_max, 0.]),
                        np.array([q0, v_max, dt]), np.array([q_fin, 0., 0.]), np.array([q_fin, 0., dt]))

    N = N0
    xg = np.stack([np.linspace(q_init, q_fin, N + 1), np.full(N + 1, v_sel), np.full(N + 1, dt)], axis=1)
    ug = np.zeros((N, 1))
    norm_old = v_max
    while True:
        ans = yield req(N, xg, ug, q_init)
        if ans.status != 0:
            raise RuntimeError("Sorry, the solver failed")
        norm_new = abs(ans.x[0, 1])
        if norm_new > norm_old + 1e-4 and N + 1 <= N_CAP:
            norm_old = norm_new
            xg, ug = ans.x.copy(), np.vstack([ans.u, np.zeros((1, 1))])
            N += 1
        else:
            break
    x_sol, u_sol = ans.x.copy(), ans.u.copy()
    rows = [x_sol[0, :2].copy()]
    v_out = x_sol[0, 1] - eps * cost_dir
    at_limit = v_out > v_max or v_out < -v_max
    for f in range(1, N):
        if at_limit:
            v_out = x_sol[f, 1] - eps * cost_dir
            if v_out > v_max or v_out < -v_max:
                rows.append(x_sol[f, :2].copy())
            else:
                norm_old = abs(x_sol[f, 1])
                N_t = N - f
                sub = yield req(N_t, x_sol[:N_t + 1], u_sol[:N_t], x_sol[f, 0])  # unshifted guess, as the driver
                if sub.status != 0:
                    raise RuntimeError("Sorry, the solver failed")
                norm_new = abs(sub.x[0, 1])
                if norm_new > norm_old + 1e-4:
                    x_sol[f:N] = sub.x[:N - f]
                    u_sol[f:N] = sub.u[:N - f]
                    v_out = sub.x[0, 1] - eps * cost_dir
                    at_limit = v_out > v_max or v_out < -v_max
                else:
                    at_limit = False
                rows.append(x_sol[f, :2].copy())
        else:
            if abs(x_sol[f, 0] - q_fin) <= 1e-3:
                break
            rows.append(x_sol[f, :2].copy())
    return rows


# ------------------------------------------------------------------------------------------------
def _gpu_backend(n, capacity, device):
    from . import engine
    solver = engine.BatchSolver(n, "vboc", capacity, N_CAP, device=device)
    return solver, (lambda n_, X, U, T: engine.sim_step(n_, X, U, T, device))


def testing_batch(n, num_prob, seed, device=0, backend=None, stats=None):
    """`Pool.map(testing, range(num_prob))` (triplependulum_testdata.py:141-142) -> X_test (rows [q, v]);
    failed problems are dropped."""
    solver, sim = backend or _gpu_backend(n, num_prob, device)
    res = run_workers(n, [testing_worker(n, _rng(seed, i)) for i in range(num_prob)], solver, sim, stats)
    return np.array([r for r in res if r is not None]).reshape(-1, 2 * n)


def data_generation_batch(n, num_prob, seed, device=0, backend=None, stats=None, restarts=True):
    """`Pool.map(data_generation, range(num_prob))` + the flattening of `traj`
    (VBOC/triplependulum_vboc.py:399-405) -> X_save (rows [q, v]).  restarts=False: the generic `VBOC/vboc.py` worker."""
    solver, sim = backend or _gpu_backend(n, num_prob, device)
    res = run_workers(n, [data_generation_worker(n, _rng(seed, i), restarts=restarts) for i in range(num_prob)], solver, sim, stats)
    rows = [np.asarray(r) for r in res if r is not None and len(r)]
    if stats is not None:
        stats["problems"] = num_prob
        stats["problems_ok"] = len(rows)
    return np.concatenate(rows).reshape(-1, 2 * n) if rows else np.empty((0, 2 * n))


def _stream_backend(n, capacity, device):
    from . import engine
    return engine.StreamSolver(n, "vboc", capacity, N_CAP, device=device)


def testing_stream(n, num_prob, seed, device=0, ssol=None, stats=None):
    """`testing_batch` through the streaming engine (no round barrier)."""
    own = ssol is None
    ssol = ssol or _stream_backend(n, num_prob, device)
    res = run_workers_stream(n, [testing_worker(n, _rng(seed, i)) for i in range(num_prob)], ssol, stats)
    if own:
        ssol.close()
    return np.array([r for r in res if r is not None]).reshape(-1, 2 * n)


def data_generation_stream(n, num_prob, seed, device=0, ssol=None, stats=None, restarts=True):
    """`data_generation_batch` through the streaming engine (no round barrier): same workers, same rows."""
    own = ssol is None
    ssol = ssol or _stream_backend(n, num_prob, device)
    res = run_workers_stream(n, [data_generation_worker(n, _rng(seed, i), restarts=restarts) for i in range(num_prob)], ssol, stats)
    if own:
        ssol.close()
    rows = [np.asarray(r) for r in res if r is not None and len(r)]
    if stats is not None:
        stats["problems"] = num_prob
        stats["problems_ok"] = len(rows)
    return np.concatenate(rows).reshape(-1, 2 * n) if rows else np.empty((0, 2 * n))


def testing_device(n, num_prob, seed, device=0, dgen=None, stats=None):
    """`testing_batch` with the per-problem state machine on the device (`vboc_testdata_run`): same seed, same X_test."""
    from . import engine
    own = dgen is None
    dgen = dgen or engine.DataGenerator(n, num_prob, device=device)
    rows, st = dgen.run_testing(testing_inputs(n, num_prob, seed))
    if stats is not None:
        stats.update(problems=num_prob, problems_ok=int((st["status"] == 0).sum()), solves=int(st["solves"].sum()),
                     converged=int(st["converged"].sum()), kernel_ms=dgen.last_kernel_ms)
    if own:
        dgen.close()
    return rows


def data_generation_device(n, num_prob, seed, device=0, dgen=None, stats=None, first=0, sharded=False):
    """`data_generation_batch` with the per-problem state machine ON THE DEVICE (`vboc_datagen_run`): one
    kernel launch, one warp per problem, the host only prepares the random draws and harvests the rows.  Same seed,
    same rows as the host generators.  sharded=True (under torch.distributed): this rank runs its contiguous shard of
    the problem range and the rows of all ranks are all-gathered (rank order = problem order)."""
    from . import engine
    lo, hi = first, first + num_prob
    if sharded:
        import torch.distributed as dist
        from . import distributed as vd
        a, b = vd.shard_range(num_prob, dist.get_rank(), dist.get_world_size())
        lo, hi = first + a, first + b
    own = dgen is None
    rows = np.empty((0, 2 * n))
    st = None
    if hi > lo:
        dgen = dgen or engine.DataGenerator(n, hi - lo, device=device)
        rows, st = dgen.run(dg_inputs(n, hi - lo, seed, first=lo))
        if own:
            dgen.close()
    if stats is not None and st is not None:
        stats["problems"] = hi - lo
        stats["problems_ok"] = int((st["status"] != 1).sum())
        stats["solves"] = int(st["solves"].sum())
        stats["converged"] = int(st["converged"].sum())
        stats["sim_steps"] = int(st["sim_steps"].sum())
        stats["kernel_ms"] = dgen.last_kernel_ms if not own else stats.get("kernel_ms")
        stats["overflow"] = int((st["status"] == 2).sum())
    if sharded:
        from . import distributed as vd
        rows = vd.all_gather_rows(rows)
    return rows


# ------------------------------------------------------------------------------------------------
# AL: feasibility labels and the active-learning query
def al_label_batch(n, X, device=0, solver=None, N=100, Tf=1.0, x_guess=None, guess_net=None):
    """Batched `testing(s0)` / `testing_guess(s0)` of the AL drivers (AL/triplependulum_al.py:24-62):
    states outside the position / velocity box are labelled unviable without solving (:27-28), the others
    get one SQP_RTI solve (`compute_problem`, AL/triplependulum_class_al.py:148-169): label 1 if status 0,
    0 if status 4 (QP failure), 2 otherwise.  `x_guess` (B, N+1, 2n) replaces the constant guess
    (`compute_problem_nnguess`, :171-201); `guess_net = (model, mean, std)` evaluates the guess network INSIDE the
    solve kernel instead (`vboc_set_guess_network`): no guess array is built or copied.  Returns labels (B,) and the
    (N+1) x 2n trajectories (NaN rows where no viable trajectory exists)."""
    from . import engine
    from ._lib import MODE_RTI
    mdl = pr.Model(n)
    X = np.asarray(X, dtype=float)
    B = X.shape[0]
    inside = (np.all(X[:, :n] >= mdl.thetamin, axis=1) & np.all(X[:, :n] <= mdl.thetamax, axis=1)
              & np.all(np.abs(X[:, n:]) <= mdl.dthetamax, axis=1))
    labels = np.zeros(B, dtype=np.int64)
    traj = np.full((B, N + 1, 2 * n), np.nan)
    idx = np.where(inside)[0]
    if idx.size:
        bp = pr.al_problems(n, X[idx], N=N, Tf=Tf, x_guess=None if x_guess is None else np.asarray(x_guess)[idx])
        own = solver is None
        sol = solver or engine.BatchSolver(n, "al", idx.size, N, device=device)
        if guess_net is not None:
            sol.set_guess_network(*guess_net)
        out = sol.solve(bp, MODE_RTI)
        if guess_net is not None and not own:
            sol.set_guess_network(None, 0.0, 1.0)
        if own:
            sol.close()
        st = out["status"]
        labels[idx] = np.where(st == 0, 1, np.where(st == 4, 0, 2))
        ok = st == 0
        traj[idx[ok]] = out["x"][ok][:, :N + 1]
    return labels, traj


def al_query(net, pool, mean, std, B, sharded=False):
    """Entropy of sigmoid(model((x - mean)/std)) over the unlabeled pool on the GPU and the B most
    uncertain samples, largest index first (AL/triplependulum_al.py:253-281).  Returns (indices into `pool`,
    entropy of `pool`, largest selected entropy).

    sharded=False: `pool` is the whole pool (single process, or every rank holds a replica and selects the same
    rows).  sharded=True (under torch.distributed): every rank holds a DIFFERENT shard of the pool; the selection
    is the global top-B (`distributed.sharded_topk`) and the returned indices are this rank's part of it, local
    to its shard -- the ranks' parts are disjoint and add up to B."""
    from . import distributed as vd
    from . import nn as vnn
    _, etp = net.entropy(pool, mean, std)
    if sharded:
        idx, emax = vd.sharded_topk(etp, B)
        return idx.tolist(), etp, emax
    idx = vnn.select_max_entropy(etp, B)
    return idx, etp, float(np.max(etp[idx])) if len(idx) else 0.0


def al_query_resident(rp, net, mean, std, B, sharded=False):
    """`al_query` on a pool that is RESIDENT on the device (`nn.ResidentPool`): entropy of every pool row
    by the fused MLP kernel, top-B by the device radix select, removal of the queried rows by the device compaction --
    only the B selected rows come back.  Returns (indices into the pool as it was before the removal, largest first;
    the selected rows (float32); largest selected entropy).
    sharded=True: every rank holds a different shard; each rank's candidates are its local top-B, the global top-B is
    cut from the gathered candidate scores and each rank keeps (and removes) its part of it."""
    rp.score(net, mean, std)
    k = min(int(B), len(rp))
    idx, x, sc = rp.select(k)
    emax = float(sc.max()) if k else 0.0
    if sharded:
        from . import distributed as vd
        allsc = vd.all_gather_rows(sc.astype(np.float64)[:, None])[:, 0]
        kk = min(int(B), allsc.shape[0])
        if kk:
            cut = np.partition(allsc, -kk)[-kk]
            emax = float(allsc.max())
            # ties at the cut are shared out in rank order so that the ranks' parts add up to exactly B
            above = int((sc > cut).sum())
            ties_all = vd.all_gather_rows(np.array([[float((sc == cut).sum())]]))[:, 0].astype(int)
            import torch.distributed as dist
            want_ties = kk - int((allsc > cut).sum())
            before = int(ties_all[:dist.get_rank()].sum())
            mine = max(0, min(int(ties_all[dist.get_rank()]), want_ties - before))
            idx, x, sc = rp.select(above + mine)
    rp.remove_selected()
    return idx, x, emax


# ------------------------------------------------------------------------------------------------
# on-disk formats read by the untouched *_comparison.py scripts: vboc_b200/io.py
def save_testdata(n, X_test, directory="."):
    """`np.save('data3_test.npy', X_test)` (triplependulum_testdata.py:144-145): rows [q, v]."""
    from . import io as vio
    return vio.save_testdata(n, X_test, directory)


def save_vboc_data(n, X_save, directory="."):
    """`np.save('data_3dof_vboc', np.asarray(X_save))` (VBOC/triplependulum_vboc.py:575-584): rows [q, v]."""
    from . import io as vio
    return vio.save_run(n, "vboc", directory, data=X_save)["data"]


def pendulum_data_generation(device=0, backend=None, stats=None):
    """The 1-DOF driver's data generation (VBOC/pendulum_vboc.py:54-223): both sides -> X_save (rows [q, v])."""
    mdl = pr.Model(1)
    solver, sim = backend or _gpu_backend(1, 2, device)
    res = run_workers(1, [pendulum_worker(-mdl.dthetamax), pendulum_worker(mdl.dthetamax)], solver, sim, stats)
    return np.concatenate([np.asarray(r) for r in res]).reshape(-1, 2)
