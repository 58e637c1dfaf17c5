"""The reference's OWN solver path, whenever it can run: the unmodified
`VBOC/triplependulum_class_vboc.py` classes on real acados under `multiprocessing.Pool(os.cpu_count())`, fed with the
same seeded problem list as the CUDA engine (vboc_b200.problems).

`available()` probes `import acados_template`, `import casadi` and a copy of the reference scripts
(`$VBOC_REFERENCE_DIR`, `baseline/_ref`, `/root/reference`).  In the build container and on the GPU boxes of this
pool the probe fails (no acados / casadi wheel, no network: DESIGN.md section 6), and every caller falls back to the
oracle port and says so -- but nothing else has to change on a machine where acados exists: `bench.py --impl
reference` then reports `cpu_baseline.kind = "acados"` and `tools/agreement.py` compares against it.

Only the `Pool.map` region is timed: solver construction / code generation happens in the pool initialiser
(the reference builds its solver objects at import time, before the Pool, VBOC/triplependulum_vboc.py:377-402).
"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = [os.environ.get("VBOC_REFERENCE_DIR"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"]

_CLASS = {2: ("doublependulum_class_vboc", "OCPdoublependulumINIT"), 3: ("triplependulum_class_vboc", "OCPtriplependulumINIT")}


def reference_dir():
    for d in _CANDIDATES:
        if d and os.path.isfile(os.path.join(d, "VBOC", "triplependulum_class_vboc.py")):
            return d
    return None


def available():
    """(ok, why).  ok only if acados_template and casadi import AND the reference scripts are present."""
    try:
        importlib.import_module("acados_template")
        importlib.import_module("casadi")
    except Exception as e:  # ModuleNotFoundError here; a broken install raises other things
        return False, f"acados not importable: {type(e).__name__}: {e}"
    d = reference_dir()
    if d is None:
        return False, "acados imports, but no copy of the reference scripts was found (set VBOC_REFERENCE_DIR)"
    return True, d


_ocp = None


def _init(ref_dir, n, workdir):
    """Pool initialiser: one unmodified reference solver object per worker process (acados writes its generated code
    into the CWD, so every worker gets its own directory)."""
    global _ocp
    d = os.path.join(workdir, f"w{os.getpid()}")
    os.makedirs(d, exist_ok=True)
    os.chdir(d)
    sys.path.insert(0, os.path.join(ref_dir, "VBOC"))
    mod, cls = _CLASS[n]
    _ocp = getattr(importlib.import_module(mod), cls)()


def _solve(args):
    """One `OCP_solve` exactly as `data_generation` calls it (VBOC/triplependulum_vboc.py:110-129)."""
    N, xg, ug, p, lbx, ubx, lbu, ubu, lbx0, ubx0, lbxN, ubxN = args
    ocp = _ocp
    ocp.N = N
    ocp.ocp_solver.set_new_time_steps(np.full((N,), 1.))
    ocp.ocp_solver.update_qp_solver_cond_N(N)
    status = ocp.OCP_solve(xg, ug, p, lbx, ubx, lbu, ubu, lbx0, ubx0, lbxN, ubxN)
    x = np.array([ocp.ocp_solver.get(i, "x") for i in range(N + 1)])
    u = np.array([ocp.ocp_solver.get(i, "u") for i in range(N)])
    return int(status), float(ocp.ocp_solver.get_cost()), x, u


def solve_batch(n, bp, processes=None, workdir="/tmp/vboc_acados_arm"):
    """Solve the batched problem dict `bp` (vboc_b200.problems.sample_vboc / sample_testdata) with the reference
    classes.  Returns dict(status, cost, x, u, wall_s, processes) shaped like `engine.BatchSolver.solve`."""
    import multiprocessing as mp
    ok, where = available()
    if not ok:
        raise RuntimeError(where)
    processes = processes or os.cpu_count()
    B = len(bp["N"])
    jobs = []
    for b in range(B):
        N = int(bp["N"][b])
        jobs.append((N, bp["x_guess"][b, :N + 1], bp["u_guess"][b, :N], bp["p"][b], bp["lbx"][b], bp["ubx"][b], bp["lbu"][b],
                     bp["ubu"][b], bp["lbx0"][b], bp["ubx0"][b], bp["lbxN"][b], bp["ubxN"][b]))
    with mp.get_context("fork").Pool(processes, initializer=_init, initargs=(where, n, workdir)) as pool:
        pool.map(_noop, range(processes))          # make sure every worker has built its solver before timing
        t0 = time.perf_counter()
        res = pool.map(_solve, jobs, chunksize=max(1, B // (8 * processes)))
        wall = time.perf_counter() - t0
    Nmax = bp["x_guess"].shape[1] - 1
    x = np.zeros((B, Nmax + 1, bp["x_guess"].shape[2]))
    u = np.zeros((B, Nmax, n))
    for b, (_, _, xb, ub) in enumerate(res):
        x[b, :xb.shape[0]], u[b, :ub.shape[0]] = xb, ub
    return dict(status=np.array([r[0] for r in res]), cost=np.array([r[1] for r in res]), x=x, u=u, wall_s=wall,
                processes=processes)


def _noop(_):
    return 0


if __name__ == "__main__":
    print(available())
