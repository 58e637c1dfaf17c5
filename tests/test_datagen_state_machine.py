"""The on-device data_generation state machine (vboc_b200/csrc/datagen_warp.h) against the host
generators (`drivers.data_generation_worker`, the faithful restatement of VBOC/triplependulum_vboc.py:19-370): the
kernel SOURCE is compiled for the host by tools/emu and must return the same rows, solve counts and simulator steps
as the generators driven with the same solver source -- i.e. the control flow (extensions, restarts, sub-OCP chains,
twin simulation, row filter) is the same.  The GPU run of the compiled kernel is in tests/test_gpu_datagen.py."""
import os
import sys

import numpy as np
import pytest

from vboc_b200 import drivers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools", "emu"))


@pytest.fixture(scope="module")
def emu():
    import emu as e
    e.build()
    return e


class EmuBackend:
    """`BatchSolver`-like backend over the host emulation of the warp solver (the same code the state machine calls)."""
    N_max = drivers.N_CAP

    def __init__(self, emu, oracle, n):
        self.emu, self.n, self.oracle = emu, n, oracle
        self.opts = emu.Opts()
        oo = oracle.default_opts(0)
        for f, _ in emu.Opts._fields_:
            setattr(self.opts, f, getattr(oo, f))

    def solve(self, bp, mode):
        return self.emu.solve_batch(self.n, 0, mode, bp, self.opts)

    def sim(self, n, X, U, T):
        return np.stack([self.oracle.rk4(n, 1, x, u, T) for x, u in zip(X, U)])


@pytest.mark.parametrize("n,num", [(3, 6), (2, 10)])
def test_state_machine_equals_host_generators(oracle, emu, n, num):
    be = EmuBackend(emu, oracle, n)
    workers = [drivers.data_generation_worker(n, drivers._rng(4, i)) for i in range(num)]
    st = {}
    ref = drivers.run_workers(n, workers, be, be.sim, st)
    out, cnt = emu.datagen_run(n, drivers.dg_inputs(n, num, 4), be.opts)
    assert sum(c["solves"] for c in cnt) == st["solves"] and sum(c["converged"] for c in cnt) == st["converged"]
    for b in range(num):
        if ref[b] is None:
            assert out[b] is None and cnt[b]["status"] == 1
            continue
        r = np.asarray(ref[b]).reshape(-1, 2 * n)
        assert out[b].shape == r.shape, (b, out[b].shape, r.shape)
        assert np.abs(out[b] - r).max() < 1e-9, (b, np.abs(out[b] - r).max())
    assert any(c["sim_steps"] > 0 for c in cnt) and any(c["solves"] > 1 for c in cnt)


def test_inputs_follow_the_generator_stream():
    """dg_inputs draws exactly what the generator draws: the first request of the generator carries the same p / bounds."""
    n = 3
    inp = drivers.dg_inputs(n, 5, seed=9, first=3)
    for b in range(5):
        req = next(drivers.data_generation_worker(n, drivers._rng(9, 3 + b)))
        assert np.array_equal(req.p, inp["p"][b]) and np.array_equal(req.q_init_lb, inp["lb0"][b])
        assert np.array_equal(req.q_init_ub, inp["ub0"][b])
        assert inp["lb0"][b, inp["joint_sel"][b]] == inp["ub0"][b, inp["joint_sel"][b]]
    assert np.abs(inp["retry"]).max() <= 0.01 and np.abs(inp["retry"]).min() > 0


@pytest.mark.parametrize("n,num", [(2, 12), (3, 8)])
def test_testdata_state_machine_equals_host_generators(oracle, emu, n, num):
    """`testing(v)` on the device (DataGen::run_testing) against `drivers.testing_worker` over the same solver source."""
    be = EmuBackend(emu, oracle, n)
    workers = [drivers.testing_worker(n, drivers._rng(2, i)) for i in range(num)]
    st = {}
    ref = drivers.run_workers(n, workers, be, be.sim, st)
    rows, cnt = emu.testdata_run(n, drivers.testing_inputs(n, num, 2), be.opts)
    assert sum(c["solves"] for c in cnt) == st["solves"] and sum(c["converged"] for c in cnt) == st["converged"]
    for b in range(num):
        if ref[b] is None:
            assert cnt[b]["status"] == 1
        else:
            assert cnt[b]["status"] == 0 and np.abs(rows[b] - ref[b]).max() < 1e-9, (b, np.abs(rows[b] - ref[b]).max())
    assert any(c["solves"] > 1 for c in cnt)
