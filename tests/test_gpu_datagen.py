"""The device-resident data_generation state machine (`vboc_datagen_run`) on the GPU: same rows as the
host generators served by the batched solver (`drivers.data_generation_batch`), hence -- through
tests/test_gpu_drivers.py -- as the generators on the oracle."""
import numpy as np
import pytest

from vboc_b200 import drivers, engine

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,num", [(3, 48), (2, 48)])
def test_device_state_machine_equals_host_generators(n, num):
    s1, s2 = {}, {}
    ref = drivers.data_generation_batch(n, num, seed=4, stats=s1)
    out = drivers.data_generation_device(n, num, seed=4, stats=s2)
    assert s1["solves"] == s2["solves"] and s1["converged"] == s2["converged"]
    assert s1["problems_ok"] == s2["problems_ok"] and s2["overflow"] == 0
    assert ref.shape == out.shape
    # two instantiations of the solver template (batched kernel / state-machine kernel): the compiler schedules their
    # FMAs differently and an SQP run amplifies that rounding (same bound as stream vs batch)
    assert np.abs(ref - out).max() < 1e-6, np.abs(ref - out).max()


def test_device_state_machine_counters_and_determinism():
    n, num = 3, 64
    dg = engine.DataGenerator(n, num)
    inp = drivers.dg_inputs(n, num, seed=7)
    rows1, st1 = dg.run(inp)
    rows2, st2 = dg.run(inp)
    dg.close()
    assert np.array_equal(rows1, rows2) and np.array_equal(st1["n_rows"], st2["n_rows"])   # bit-identical re-run
    assert rows1.shape[0] == st1["n_rows"].sum() and (st1["status"] != 2).all()
    assert (st1["solves"] >= st1["converged"]).all() and (st1["solves"] >= 1).all() and (st1["solves"] <= 10 + 5 * 128).all()
    ok = st1["status"] == 0
    assert ok.mean() > 0.9 and (st1["n_rows"][ok] >= 1).all() and (st1["n_rows"][~ok] == 0).all()
    # rows are states inside the box
    from vboc_b200 import problems as pr
    mdl = pr.Model(n)
    assert (rows1[:, :n] >= mdl.thetamin - 1e-9).all() and (rows1[:, :n] <= mdl.thetamax + 1e-9).all()
    assert (np.abs(rows1[:, n:]) <= mdl.dthetamax + 1e-6).all()


def test_datagen_refusals():
    from vboc_b200._lib import VbocError
    dg = engine.DataGenerator(3, 4)
    inp = drivers.dg_inputs(3, 4, seed=1)
    bad = dict(inp)
    bad["ub0"] = inp["ub0"].copy()
    bad["ub0"][:, 6] = 2e-2          # dt not pinned
    with pytest.raises(VbocError):
        dg.run(bad)
    with pytest.raises(VbocError):
        dg.run(drivers.dg_inputs(3, 8, seed=1))   # more problems than the capacity
    dg.close()
    with pytest.raises(VbocError):
        engine.DataGenerator(1, 4)   # the 1-DOF driver has its own (free-dt) flow


@pytest.mark.parametrize("n,num", [(2, 64), (3, 32)])
def test_device_testdata_state_machine_equals_host_generators(n, num):
    """`vboc_testdata_run` (`testing(v)`, triplependulum_testdata.py:9-125, as a state machine on the device) against the
    host generators served by the batched solver: same solve counts, same X_test."""
    s1, s2 = {}, {}
    ref = drivers.testing_batch(n, num, seed=2, stats=s1)
    out = drivers.testing_device(n, num, seed=2, stats=s2)
    assert s1["solves"] == s2["solves"] and s1["converged"] == s2["converged"]
    assert ref.shape == out.shape and s2["problems_ok"] == ref.shape[0]
    assert np.abs(ref - out).max() < 1e-6, np.abs(ref - out).max()
