"""GPU inference of the reference's MLPs (my_nn.py) for the NN filter: viability labels / margins
(VBOC) and entropy-based active-learning queries (AL).  Training stays in PyTorch (north_star)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class MLP:
    """3-layer MLP on the device.  `MLP.from_torch(model)` takes a `NeuralNetDIR` / `NeuralNetCLS`
    (anything with a `linear_relu_stack` Sequential of Linear / ReLU modules)."""

    def __init__(self, W1, b1, W2, b2, W3, b3, final_relu, device=0):
        f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        W1, b1, W2, b2, W3, b3 = map(f, (W1, b1, W2, b2, W3, b3))
        self.n_in, self.hidden, self.n_out = W1.shape[1], W1.shape[0], W3.shape[0]
        assert W2.shape == (self.hidden, self.hidden) and W3.shape[1] == self.hidden
        self._h = C.c_void_p()
        check(_lib.lib().vboc_mlp_create(int(device), self.n_in, self.hidden, self.n_out, int(bool(final_relu)),
                                         _fp(W1), _fp(b1), _fp(W2), _fp(b2), _fp(W3), _fp(b3), C.byref(self._h)))

    @classmethod
    def from_torch(cls, model, device=0):
        mods = list(model.linear_relu_stack)
        lin = [m for m in mods if hasattr(m, "weight")]
        assert len(lin) == 3, "expected Linear-ReLU-Linear-ReLU-Linear[-ReLU]"
        final_relu = not hasattr(mods[-1], "weight")
        g = lambda t: t.detach().cpu().numpy()
        return cls(g(lin[0].weight), g(lin[0].bias), g(lin[1].weight), g(lin[1].bias), g(lin[2].weight),
                   g(lin[2].bias), final_relu, device)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value and _lib is not None:
            _lib.lib().vboc_mlp_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def _run(self, X, mode, mean=0.0, std=1.0, safety_margin=0.0, want_label=False):
        X = np.ascontiguousarray(X, dtype=np.float32)
        B = X.shape[0]
        out = np.empty((B, self.n_out), dtype=np.float32)
        aux = np.empty(B, dtype=np.float32)
        lab = np.empty(B, dtype=np.int32) if want_label else None
        check(_lib.lib().vboc_mlp_forward(self._h, B, _fp(X), mode, float(mean), float(std), float(safety_margin),
                                          _fp(out), _fp(aux),
                                          lab.ctypes.data_as(C.POINTER(C.c_int)) if want_label else None))
        return out, aux, lab

    @property
    def last_kernel_ms(self):
        """Device time of the last forward kernel (CUDA events around the launch, no copies)."""
        return _lib.lib().vboc_mlp_last_kernel_ms(self._h)

    def forward(self, X):
        """net(X) for already normalised inputs."""
        return self._run(X, 0)[0]

    def viability(self, X, mean, std, safety_margin=0.0):
        """VBOC filter (VBOC/triplependulum_vboc.py:604-620): phi, label (1 viable, 0 not), margin."""
        out, aux, lab = self._run(X, 1, mean, std, safety_margin, want_label=True)
        return out[:, 0], lab, aux

    def entropy(self, X, mean, std):
        """AL query score (AL/triplependulum_al.py:253-264): logits and entropy of sigmoid(logits)."""
        out, aux, _ = self._run(X, 2, mean, std)
        return out, aux


def select_max_entropy(entropy, B):
    """Indices of the B most uncertain samples, largest index first (AL/triplependulum_al.py:267-270);
    across ranks use `vboc_b200.distributed.global_topk`."""
    idx = np.argpartition(entropy, -B)[-B:].tolist()
    idx.sort(reverse=True)
    return idx


class ResidentPool:
    """The AL drivers' unlabeled pool kept in HBM across rounds (`vboc_pool_*`): `score(net, mean, std)` runs the fused
    MLP + entropy kernel over the resident rows, `select(k)` is the device top-k (indices largest first, rows, scores),
    `remove_selected()` the device `np.delete`.  Per round only the k selected rows cross the host link."""

    def __init__(self, X, device=0, capacity=None):
        X = np.ascontiguousarray(X, dtype=np.float32)
        self.n_in = X.shape[1]
        self._h = C.c_void_p()
        check(_lib.lib().vboc_pool_create(int(device), self.n_in, int(capacity or max(len(X), 1)), C.byref(self._h)))
        check(_lib.lib().vboc_pool_upload(self._h, len(X), _fp(X)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value and _lib is not None:
            _lib.lib().vboc_pool_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __len__(self):
        return int(_lib.lib().vboc_pool_size(self._h))

    def score(self, net, mean, std):
        check(_lib.lib().vboc_pool_score(self._h, net._h, float(mean), float(std)))
        return _lib.lib().vboc_pool_last_score_ms(self._h)

    def select(self, k):
        k = int(k)
        idx = np.empty(k, dtype=np.int64)
        x = np.empty((k, self.n_in), dtype=np.float32)
        sc = np.empty(k, dtype=np.float32)
        check(_lib.lib().vboc_pool_select(self._h, k, idx.ctypes.data_as(C.POINTER(C.c_longlong)), _fp(x), _fp(sc)))
        return idx, x, sc

    def remove_selected(self):
        check(_lib.lib().vboc_pool_remove_selected(self._h))

    def scores(self):
        sc = np.empty(len(self), dtype=np.float32)
        check(_lib.lib().vboc_pool_download_scores(self._h, _fp(sc)))
        return sc

    def rows(self):
        x = np.empty((len(self), self.n_in), dtype=np.float32)
        check(_lib.lib().vboc_pool_download(self._h, _fp(x)))
        return x
