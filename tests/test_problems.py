"""Seeded problem sampler (the reference draws with unseeded `random`)."""
import numpy as np

from vboc_b200 import problems as pr


def test_vboc_sampler_invariants():
    for n in (2, 3):
        bp = pr.sample_vboc(n, 64, seed=3)
        d = bp["p"][:, :n]
        assert np.allclose(np.linalg.norm(d, axis=1), 1.0)
        assert (bp["p"][:, n] == 0).all()
        C0 = bp["C0"]
        assert np.allclose(C0[:, :, n:2 * n], np.eye(n) - d[:, :, None] * d[:, None, :])
        assert (C0[:, :, :n] == 0).all() and (C0[:, :, 2 * n] == 0).all()
        rows = np.arange(64)
        js = bp["joint_sel"]
        assert (bp["lbx0"][rows, js] == bp["ubx0"][rows, js]).all()
        assert np.sign(d[rows, js]).tolist() == bp["vel_sel"].tolist()
        assert (bp["lbxN"][:, n:2 * n] == 0).all() and (bp["ubxN"][:, n:2 * n] == 0).all()
        assert (bp["x_guess"][:, :, 2 * n] == 1e-2).all()


def test_streams_do_not_depend_on_batch_size():
    a = pr.sample_vboc(3, 8, seed=5)
    b = pr.sample_vboc(3, 32, seed=5)
    for k in ("p", "lbx0", "x_guess"):
        assert np.array_equal(a[k], b[k][:8])


def test_al_sampler():
    bp = pr.sample_al(3, 100, seed=1)
    assert np.array_equal(bp["lbx0"], bp["ubx0"])
    assert (np.abs(bp["x0"][:, 3:]) <= 10.5).all()
    assert (bp["x_guess"][:, :, 3:] == 0).all()
