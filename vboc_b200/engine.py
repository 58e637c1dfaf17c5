"""Batched OCP engine: thin Python face of the C-ABI (one handle = one GPU, one system, one family).

`BatchSolver.solve(bp)` is the batched equivalent of calling the reference's `OCP_solve(...)` /
`compute_problem(...)` once per problem under `multiprocessing.Pool` (VBOC/triplependulum_vboc.py:399-402,
AL/triplependulum_al.py:132-134): `bp` holds the per-problem arrays those methods would pass stage by
stage to acados (`vboc_b200.problems` builds them).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import FAMILY_AL, FAMILY_MPC, FAMILY_VBOC, MODE_RTI, MODE_SQP, Opts, Stats, check

_FAM = {"vboc": FAMILY_VBOC, "al": FAMILY_AL, "mpc": FAMILY_MPC}
_STAT_FIELDS = ("status", "sqp_iter", "qp_iter", "ls_evals", "qp_status", "cost",
                "res_stat", "res_eq", "res_ineq", "res_comp")
_STATS_DTYPE = np.dtype([("status", "i4"), ("sqp_iter", "i4"), ("qp_iter", "i4"), ("ls_evals", "i4"),
                         ("qp_status", "i4"), ("pad_", "i4"), ("cost", "f8"), ("res_stat", "f8"),
                         ("res_eq", "f8"), ("res_ineq", "f8"), ("res_comp", "f8")])
assert _STATS_DTYPE.itemsize == C.sizeof(Stats)


def default_opts(family):
    o = Opts()
    _lib.lib().vboc_default_opts(_FAM.get(family, family), C.byref(o))
    return o


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _c(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


class BatchSolver:
    def __init__(self, n, family, batch_capacity, N_max, device=0, opts=None):
        self.n, self.family = int(n), _FAM.get(family, family)
        self.cap, self.N_max, self.device = int(batch_capacity), int(N_max), int(device)
        self.nx = 2 * self.n + (self.family == FAMILY_VBOC)
        self.nu = self.n
        self._h = C.c_void_p()
        check(_lib.lib().vboc_create(self.n, self.family, self.cap, self.N_max, self.device, C.byref(self._h)))
        self.opts = opts or default_opts(self.family)
        self.set_opts(self.opts)
        self._batch = 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value and _lib is not None:
            _lib.lib().vboc_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def set_opts(self, opts):
        self.opts = opts
        check(_lib.lib().vboc_set_opts(self._h, C.byref(opts)))

    def set_stream(self, cuda_stream_ptr):
        check(_lib.lib().vboc_set_stream(self._h, C.c_void_p(int(cuda_stream_ptr))))

    # -- data marshalling -------------------------------------------------------------------
    def _pack(self, bp):
        Nv = np.ascontiguousarray(bp["N"], dtype=np.int32)
        B = Nv.shape[0]
        xg, ug = _c(bp["x_guess"]), _c(bp["u_guess"])
        if xg.shape[1] != self.N_max + 1:  # re-stride to the solver's N_max
            x2 = np.zeros((B, self.N_max + 1, self.nx))
            u2 = np.zeros((B, self.N_max, self.nu))
            m = min(xg.shape[1], self.N_max + 1)
            x2[:, :m] = xg[:, :m]
            u2[:, :m - 1] = ug[:, :m - 1]
            xg, ug = x2, u2
        assert xg.shape == (B, self.N_max + 1, self.nx) and ug.shape == (B, self.N_max, self.nu)
        arrs = [xg, ug] + [_c(bp.get(k)) for k in
                           ("p", "lbx0", "ubx0", "lbx", "ubx", "lbxN", "ubxN", "lbu", "ubu", "C0")]
        return B, Nv, arrs, float(bp.get("Tf", 1.0))

    def upload(self, bp):
        B, Nv, arrs, Tf = self._pack(bp)
        check(_lib.lib().vboc_upload(self._h, B, Nv.ctypes.data_as(C.POINTER(C.c_int)),
                                     *[_dp(a) for a in arrs], Tf))
        self._batch = B

    def solve_resident(self, mode=MODE_SQP):
        check(_lib.lib().vboc_solve_resident(self._h, int(mode)))
        return _lib.lib().vboc_last_kernel_ms(self._h)

    def solve_resident_async(self, mode=MODE_SQP):
        check(_lib.lib().vboc_solve_resident_async(self._h, int(mode)))

    def sync(self):
        check(_lib.lib().vboc_sync(self._h))
        return _lib.lib().vboc_last_kernel_ms(self._h)

    def download(self):
        B = self._batch
        x = np.empty((B, self.N_max + 1, self.nx))
        u = np.empty((B, self.N_max, self.nu))
        st = np.empty(B, dtype=_STATS_DTYPE)
        check(_lib.lib().vboc_download(self._h, _dp(x), _dp(u), st.ctypes.data_as(C.POINTER(Stats))))
        return self._result(x, u, st)

    # -- MPC family ------------------------------------------------------------
    def set_mpc(self, net, mean, std, safety_margin, W, W_e, lh=0.0, uh=1e6, vstart=None):
        """The margin network (a torch module with `linear_relu_stack`, or a dict of W1, b1, W2, b2, W3, b3 arrays in
        nn.Linear layout), its normalisation, the constraint's bounds and the diagonals of cost.W ([x; u] order) /
        cost.W_e."""
        if not isinstance(net, dict):
            lin = [m for m in net.linear_relu_stack if hasattr(m, "weight")]
            g = lambda t: t.detach().cpu().numpy()
            net = dict(W1=g(lin[0].weight), b1=g(lin[0].bias), W2=g(lin[1].weight), b2=g(lin[1].bias),
                       W3=g(lin[2].weight), b3=g(lin[2].bias))
        f = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float32).ravel())
        w = [f(net[k]) for k in ("W1", "b1", "W2", "b2", "W3", "b3")]
        H = w[1].shape[0]
        assert w[0].shape[0] == H * 2 * self.n and w[2].shape[0] == H * H and w[4].shape[0] == H and w[5].shape[0] == 1
        fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
        check(_lib.lib().vboc_set_mpc(self._h, H, *[fp(a) for a in w], float(mean), float(std), float(safety_margin),
                                      float(lh), float(uh), _dp(_c(W)), _dp(_c(W_e))))
        if vstart is not None:   # vel_norm = |x[vstart:]| (the triple-pendulum classes' x[2:])
            check(_lib.lib().vboc_set_mpc_velnorm_start(self._h, int(vstart)))

    def set_mpc_reference(self, yref, yref_e):
        yref, yref_e = _c(np.atleast_2d(yref)), _c(np.atleast_2d(yref_e))
        check(_lib.lib().vboc_set_mpc_reference(self._h, yref.shape[0], _dp(yref), _dp(yref_e)))

    def set_mpc_rows(self, Z):
        """Soft margin rows at every stage (the parallel / receding / soft_traj Safe-MPC variants): Z (B, N_max+1, 4) =
        per problem and stage (Zl, Zu, zl, zu), or (B, N_max+1) = Zl only (the reference sets nothing else,
        VBOC/Safe MPC/parallel/2dof_sym.py:44-50).  None: back to the hard terminal row."""
        if Z is None:
            check(_lib.lib().vboc_set_mpc_rows(self._h, 0, None))
            return
        Z = np.asarray(Z, dtype=np.float64)
        if Z.ndim == 2:
            Z = np.concatenate([Z[:, :, None], np.zeros(Z.shape + (3,))], axis=2)
        assert Z.shape[1:] == (self.N_max + 1, 4), Z.shape
        Z = _c(Z)
        check(_lib.lib().vboc_set_mpc_rows(self._h, Z.shape[0], _dp(Z)))

    def mpc_rows(self):
        """(lam_l, lam_u, lam_sl, lam_su, sl, su) of every stage's margin row at the returned iterate, (B, N_max+1, 6)."""
        rows = np.zeros((self._batch, self.N_max + 1, 6))
        check(_lib.lib().vboc_download_mpc_rows(self._h, _dp(rows)))
        return rows

    def set_cartesian(self, xc=0.0, yc=-1.2, radius=0.2, uh=1e6, on=True):
        """VBOC family, n = 2: the end effector stays outside the circle of `radius` around (xc, yc) at stages 0..N-1
        (VBOC/Cartesian constraints/doublependulum_class_fixedveldir.py:150-158: x_c = 0, y_c = -l1 - l2/2, radius = l2/4).
        Row multipliers of a solve: mpc_rows()[..., 0:2]."""
        check(_lib.lib().vboc_set_cartesian(self._h, int(on), float(xc), float(yc), float(radius) ** 2, float(uh)))

    def mpc_multipliers(self):
        lamg = np.zeros((self._batch, 2))
        check(_lib.lib().vboc_download_mpc_multipliers(self._h, _dp(lamg)))
        return lamg

    # -- AL family: the guess network inside the kernel (compute_problem_nnguess) -----------------
    def set_guess_network(self, model, mean, std):
        """`model`: torch module with `linear_relu_stack` (2n-H-H-(N 2n)), or None to switch back to host guesses."""
        fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
        if model is None:
            z = np.zeros(1, dtype=np.float32)
            check(_lib.lib().vboc_set_guess_network(self._h, 0, 0, fp(z), fp(z), fp(z), fp(z), fp(z), fp(z), 0.0, 1.0))
            return
        lin = [m for m in model.linear_relu_stack if hasattr(m, "weight")]
        g = lambda t: np.ascontiguousarray(t.detach().cpu().numpy().astype(np.float32).ravel())
        w = [g(lin[0].weight), g(lin[0].bias), g(lin[1].weight), g(lin[1].bias), g(lin[2].weight), g(lin[2].bias)]
        check(_lib.lib().vboc_set_guess_network(self._h, w[1].shape[0], w[5].shape[0], *[fp(a) for a in w],
                                                float(mean), float(std)))

    def guess(self):
        xg = np.zeros((self._batch, self.N_max + 1, self.nx))
        check(_lib.lib().vboc_download_guess(self._h, _dp(xg)))
        return xg

    def export_multipliers(self, on=True):
        """Have the next solves keep the KKT multipliers of the returned iterates (`multipliers()`)."""
        check(_lib.lib().vboc_export_multipliers(self._h, int(bool(on))))

    def multipliers(self):
        """pi (B, N_max, 2n), lam (B, N_max+1, 3n, 2) of the last solve, see include/vboc_b200.h."""
        B, n = self._batch, self.n
        pi = np.zeros((B, self.N_max, 2 * n))
        lam = np.zeros((B, self.N_max + 1, 3 * n, 2))
        check(_lib.lib().vboc_download_multipliers(self._h, _dp(pi), _dp(lam)))
        return pi, lam

    def solve(self, bp, mode=MODE_SQP):
        """upload + solve + download in one C call (host buffers in, host buffers out)."""
        B, Nv, arrs, Tf = self._pack(bp)
        x = np.empty((B, self.N_max + 1, self.nx))
        u = np.empty((B, self.N_max, self.nu))
        st = np.empty(B, dtype=_STATS_DTYPE)
        check(_lib.lib().vboc_solve_batch(self._h, int(mode), B, Nv.ctypes.data_as(C.POINTER(C.c_int)),
                                          *[_dp(a) for a in arrs], Tf, _dp(x), _dp(u),
                                          st.ctypes.data_as(C.POINTER(Stats))))
        self._batch = B
        return self._result(x, u, st)

    @staticmethod
    def _result(x, u, st):
        out = dict(x=x, u=u)
        for f in _STAT_FIELDS:
            out[f] = st[f].copy()
        out["res"] = np.stack([st["res_stat"], st["res_eq"], st["res_ineq"], st["res_comp"]], axis=1)
        return out

    @property
    def last_kernel_ms(self):
        return _lib.lib().vboc_last_kernel_ms(self._h)


class StreamSolver:
    """Ticket queue over the same kernels (`vboc_stream_*`): `submit` launches and returns at once, `poll`
    returns the tickets whose solve has finished, `fetch` hands out one result and frees its slot.  This is
    what serves the drivers' per-problem loops (`drivers.run_workers_stream`): every worker resubmits as soon
    as ITS solve is back, instead of every round waiting for the slowest solve of the round."""

    def __init__(self, n, family, capacity, N_max, device=0, opts=None):
        self.n, self.family = int(n), _FAM.get(family, family)
        self.cap, self.N_max, self.device = int(capacity), int(N_max), int(device)
        self.nx = 2 * self.n + (self.family == FAMILY_VBOC)
        self.nu = self.n
        self._h = C.c_void_p()
        check(_lib.lib().vboc_stream_create(self.n, self.family, self.cap, self.N_max, self.device, C.byref(self._h)))
        self.opts = opts or default_opts(self.family)
        check(_lib.lib().vboc_stream_set_opts(self._h, C.byref(self.opts)))
        self._tick = np.empty(self.cap, dtype=np.int32)
        self._N = {}

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value and _lib is not None:
            _lib.lib().vboc_stream_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def set_opts(self, opts):
        self.opts = opts
        check(_lib.lib().vboc_stream_set_opts(self._h, C.byref(opts)))

    @property
    def free_slots(self):
        return _lib.lib().vboc_stream_free_slots(self._h)

    @property
    def pending(self):
        return _lib.lib().vboc_stream_pending(self._h)

    def submit(self, bp, mode=MODE_SQP):
        """bp as for `BatchSolver.solve` (x_guess / u_guess with N_max + 1 / N_max rows); returns the tickets."""
        Nv = np.ascontiguousarray(bp["N"], dtype=np.int32)
        B = Nv.shape[0]
        xg, ug = _c(bp["x_guess"]), _c(bp["u_guess"])
        assert xg.shape == (B, self.N_max + 1, self.nx) and ug.shape == (B, self.N_max, self.nu)
        arrs = [xg, ug] + [_c(bp.get(k)) for k in
                           ("p", "lbx0", "ubx0", "lbx", "ubx", "lbxN", "ubxN", "lbu", "ubu", "C0")]
        tk = np.empty(B, dtype=np.int32)
        check(_lib.lib().vboc_stream_submit(self._h, int(mode), B, Nv.ctypes.data_as(C.POINTER(C.c_int)),
                                            *[_dp(a) for a in arrs], float(bp.get("Tf", 1.0)),
                                            tk.ctypes.data_as(C.POINTER(C.c_int))))
        for t, Nb in zip(tk.tolist(), Nv.tolist()):
            self._N[t] = Nb
        return tk

    def poll(self):
        k = _lib.lib().vboc_stream_poll(self._h, self.cap, self._tick.ctypes.data_as(C.POINTER(C.c_int)))
        if k < 0:
            check(k)
        return self._tick[:k].tolist()

    def fetch(self, ticket):
        """Result of a finished ticket: dict(status, cost, x (N+1, nx), u (N, nu), sqp_iter, qp_iter, ...)."""
        Nb = self._N.pop(int(ticket))
        x = np.empty((Nb + 1, self.nx))
        u = np.empty((Nb, self.nu))
        st = np.empty(1, dtype=_STATS_DTYPE)
        check(_lib.lib().vboc_stream_fetch(self._h, int(ticket), _dp(x), _dp(u), st.ctypes.data_as(C.POINTER(Stats))))
        out = dict(x=x, u=u)
        for f in _STAT_FIELDS:
            out[f] = st[f][0].item()
        return out

    def sim_step(self, x, u, T):
        x, u = _c(np.atleast_2d(x)), _c(np.atleast_2d(u))
        xn = np.empty_like(x)
        check(_lib.lib().vboc_stream_sim_step(self._h, x.shape[0], _dp(x), _dp(u), float(T), _dp(xn)))
        return xn

    def solve(self, bp, mode=MODE_SQP):
        """Convenience (tests): submit a batch and wait for all of it; results in submission order."""
        import time
        tk = self.submit(bp, mode).tolist()
        got = {}
        while len(got) < len(tk):
            done = self.poll()
            for t in done:
                got[t] = self.fetch(t)
            if not done:
                time.sleep(1e-4)
        return [got[t] for t in tk]


_DG_DTYPE = np.dtype([(f, "i4") for f in ("status", "n_rows", "solves", "converged", "sim_steps", "sqp_iter", "qp_iter", "t_done_us")])


class DataGenerator:
    """`vboc_datagen_*`: the whole `data_generation(v)` of a batch of problems as one kernel launch (one warp per problem,
    csrc/datagen_warp.h).  `run(inputs)` takes `drivers.dg_inputs(...)` and returns (rows (total, 2n) in problem order,
    per-problem counters as a structured array)."""

    def __init__(self, n, capacity, device=0, opts=None):
        self.n, self.cap, self.device = int(n), int(capacity), int(device)
        self._h = C.c_void_p()
        check(_lib.lib().vboc_datagen_create(self.n, self.cap, self.device, C.byref(self._h)))
        if opts is not None:
            check(_lib.lib().vboc_datagen_set_opts(self._h, C.byref(opts)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value and _lib is not None:
            _lib.lib().vboc_datagen_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    @property
    def last_kernel_ms(self):
        return _lib.lib().vboc_datagen_last_kernel_ms(self._h)

    def run(self, inp, N0=100, dt=1e-2, tol=1e-3):
        B, n = len(inp["joint_sel"]), self.n
        js = np.ascontiguousarray(inp["joint_sel"], dtype=np.int32)
        arrs = [_c(inp[k]) for k in ("p", "lb0", "ub0", "retry")]
        assert arrs[0].shape == (B, n + 1) and arrs[1].shape == (B, 2 * n + 1) and arrs[3].shape == (B, 10, n + 1)
        cap_rows = B * _lib.DG_ROWS_MAX
        rows = np.empty((cap_rows, 2 * n))
        st = np.zeros(B, dtype=_DG_DTYPE)
        total = C.c_longlong(0)
        check(_lib.lib().vboc_datagen_run(self._h, B, int(N0), float(dt), float(tol), js.ctypes.data_as(C.POINTER(C.c_int)),
                                          *[_dp(a) for a in arrs], _dp(rows), cap_rows, C.byref(total),
                                          st.ctypes.data_as(C.POINTER(_lib.DgStats))))
        return rows[:total.value].copy(), st


    def run_testing(self, inp, N0=100, dt=1e-2, max_solves=60):
        """`vboc_testdata_run`: inp = drivers.testing_inputs(...); returns (X_test rows of the problems that produced one,
        per-problem counters)."""
        B, n = len(inp["ran"]), self.n
        arrs = [_c(inp[k]) for k in ("ran", "q_init", "retry")]
        assert arrs[0].shape == (B, n) and arrs[2].shape == (B, 60, 2 * n)
        rows = np.zeros((B, 2 * n))
        st = np.zeros(B, dtype=_DG_DTYPE)
        check(_lib.lib().vboc_testdata_run(self._h, B, int(N0), float(dt), int(max_solves), *[_dp(a) for a in arrs],
                                           _dp(rows), st.ctypes.data_as(C.POINTER(_lib.DgStats))))
        return rows[st["status"] == 0], st


def sim_step(n, x, u, T, device=0):
    """Batched RK4 step of the unscaled model (the reference's `sim.acados_integrator`)."""
    x, u = _c(np.atleast_2d(x)), _c(np.atleast_2d(u))
    xn = np.empty_like(x)
    check(_lib.lib().vboc_sim_step(int(n), int(device), x.shape[0], _dp(x), _dp(u), float(T), _dp(xn)))
    return xn


def fp64_peak_tflops(device=0):
    """Measured DFMA peak of the device (roofline denominator)."""
    v = C.c_double(0.0)
    check(_lib.lib().vboc_fp64_peak(int(device), C.byref(v)))
    return v.value
