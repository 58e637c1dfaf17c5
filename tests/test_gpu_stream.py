"""Streaming engine (`vboc_stream_*`, engine.StreamSolver) on the GPU: the same kernels as the batched call,
so results must be IDENTICAL to `BatchSolver.solve` on the same problems, whatever the order of completion;
the event-loop drivers must return the rows of the round-synchronous drivers."""
import numpy as np
import pytest

from vboc_b200 import drivers, engine, problems as pr
from vboc_b200._lib import MODE_RTI, MODE_SQP, ERR_ARG, ERR_UNSUPPORTED, VbocError

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [2, 3])
def test_stream_equals_batch_vboc(n):
    B = 48
    bp = pr.sample_vboc(n, B, seed=31)
    ref_solver = engine.BatchSolver(n, "vboc", B, 100)
    ref = ref_solver.solve(bp)
    ref_solver.close()
    ss = engine.StreamSolver(n, "vboc", 64, 100)
    # three launches of different sizes in flight at once
    parts = [slice(0, 5), slice(5, 30), slice(30, B)]
    tickets = []
    for sl in parts:
        sub = {k: (v[sl] if isinstance(v, np.ndarray) else v) for k, v in bp.items()}
        tickets += ss.submit(sub, MODE_SQP).tolist()
    assert ss.pending == B and ss.free_slots == 64 - B
    got = {}
    while len(got) < B:
        for t in ss.poll():
            got[t] = ss.fetch(t)
    assert ss.pending == 0 and ss.free_slots == 64
    for b, t in enumerate(tickets):
        r = got[t]
        N = int(bp["N"][b])
        assert r["status"] == ref["status"][b] and r["sqp_iter"] == ref["sqp_iter"][b]
        assert r["qp_iter"] == ref["qp_iter"][b]
        # two instantiations of the same solver template: the compiler schedules their FMAs differently, and an SQP
        # run of tens of iterations at tol_stat 1e-3 amplifies that rounding (same bound as GPU vs oracle)
        if r["status"] == 0:
            ex, eu = np.abs(r["x"] - ref["x"][b, :N + 1]).max(), np.abs(r["u"] - ref["u"][b, :N]).max()
            assert ex < 1e-5 and eu < 1e-3, (b, ex, eu)
            assert abs(r["cost"] - ref["cost"][b]) < 1e-6
    # the same problems again, submitted in one launch: slot / warp assignment differs, results must not
    again = ss.solve(bp, MODE_SQP)
    for b, t in enumerate(tickets):
        assert np.array_equal(again[b]["x"], got[t]["x"]) and np.array_equal(again[b]["u"], got[t]["u"])
        assert again[b]["qp_iter"] == got[t]["qp_iter"]
    ss.close()


def test_stream_slots_are_reused_and_al_rti(oracle):
    n = 3
    B, cap = 2048, 256
    bp = pr.sample_al(n, B, seed=5)
    ss = engine.StreamSolver(n, "al", cap, 100)
    ref = oracle.solve_batch(n, oracle.FAMILY_AL, oracle.MODE_RTI, bp)
    labels = []
    for lo in range(0, B, cap):  # capacity 256: the slots go round eight times
        sub = {k: (v[lo:lo + cap] if isinstance(v, np.ndarray) else v) for k, v in bp.items()}
        labels += [r["status"] for r in ss.solve(sub, MODE_RTI)]
    assert (np.array(labels) == ref["status"]).mean() >= 0.999
    ss.close()


def test_stream_refusals():
    n = 3
    bp = pr.sample_vboc(n, 4, seed=1)
    ss = engine.StreamSolver(n, "vboc", 2, 100)
    with pytest.raises(VbocError) as e:
        ss.submit(bp)  # more problems than free slots
    assert e.value.code == ERR_ARG
    bad = {k: (v[:2].copy() if isinstance(v, np.ndarray) else v) for k, v in bp.items()}
    bad["ubx"][:, 2 * n] *= 2.0  # dt no longer pinned
    with pytest.raises(VbocError) as e:
        ss.submit(bad)
    assert e.value.code == ERR_UNSUPPORTED
    assert ss.free_slots == 2  # nothing was taken
    ss.close()


def test_stream_sim_step_equals_sim_step():
    n = 3
    rng = np.random.default_rng(0)
    X, U = rng.normal(size=(70, 2 * n)) + np.pi, rng.normal(size=(70, n))
    ss = engine.StreamSolver(n, "vboc", 128, 100)
    assert np.array_equal(ss.sim_step(X, U, 1e-2), engine.sim_step(n, X, U, 1e-2))
    ss.close()


def test_stream_drivers_equal_round_drivers():
    n = 3
    s1, s2 = {}, {}
    a = drivers.data_generation_batch(n, 24, seed=4, stats=s1)
    b = drivers.data_generation_stream(n, 24, seed=4, stats=s2)
    assert s1["solves"] == s2["solves"] and s1["converged"] == s2["converged"]
    assert a.shape == b.shape and np.abs(a - b).max() < 1e-6
    c = drivers.testing_batch(2, 16, seed=2)
    d = drivers.testing_stream(2, 16, seed=2)
    assert c.shape == d.shape and np.abs(c - d).max() < 1e-6
