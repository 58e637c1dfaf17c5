"""Import the reference's model classes in THIS container without acados/casadi.

The reference's per-system class files (`/root/reference/VBOC/*_class_vboc.py`,
`/root/reference/AL/*_class_al.py`) build their dynamics as CasADi SX expressions and
hand them to acados.  Neither package is installed here, but the expressions are plain
Python arithmetic on symbols, so a tiny stand-in for the two modules backed by sympy is
enough to import the files *unmodified* and read `model.f_expl_expr` back as a sympy
matrix.  This is only used by `tools/make_golden.py` (build container only; the GPU box
has no /root/reference) to pin the oracle's dynamics against the reference's own text.

Nothing in here is product code.
"""
import importlib.util
import sys
import types

import sympy as sp

REF = "/root/reference"


class _Bag:
    """Attribute bag standing in for AcadosOcp / AcadosModel / AcadosSim and their sub-objects."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        b = _Bag()
        object.__setattr__(self, name, b)
        return b


class _Matrix(sp.Matrix):
    def size(self):
        return self.shape


def _install_stubs():
    casadi = types.ModuleType("casadi")

    class SX:
        @staticmethod
        def sym(name, *shape):
            return sp.Symbol(name, real=True)

    def vertcat(*args):
        return _Matrix([sp.sympify(a) for a in args])

    casadi.SX = SX
    casadi.vertcat = vertcat
    casadi.sin = sp.sin
    casadi.cos = sp.cos
    casadi.exp = sp.exp
    casadi.tanh = sp.tanh
    casadi.sqrt = sp.sqrt
    casadi.fabs = sp.Abs
    casadi.horzcat = lambda *a: _Matrix([list(a)])
    casadi.fmax = sp.Max
    casadi.norm_2 = lambda v: sp.sqrt(sum(x * x for x in v))
    casadi.dot = lambda a, b: sum(x * y for x, y in zip(a, b))
    casadi.MX = SX
    casadi.Function = _Bag
    sys.modules["casadi"] = casadi

    at = types.ModuleType("acados_template")
    for n in ("AcadosOcp", "AcadosOcpSolver", "AcadosSim", "AcadosSimSolver", "AcadosModel"):
        setattr(at, n, type(n, (_Bag,), {}))
    sys.modules["acados_template"] = at


def load(relpath, modname):
    """Import /root/reference/<relpath> under the stubs and return the module."""
    _install_stubs()
    spec = importlib.util.spec_from_file_location(modname, f"{REF}/{relpath}")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def dynamics_exprs():
    """Return {name: (f_expl sympy Matrix, x symbols, u symbols)} for the six model classes
    the hot path uses (VBOC time-scaled + AL plain)."""
    out = {}
    m = load("VBOC/pendulum_class_vboc.py", "ref_p1_vboc")
    o = m.OCPpendulum.__new__(m.OCPpendulum)
    # The 1-DOF class builds its solver inside __init__; AcadosOcpSolver is a stub so this is safe.
    m.OCPpendulum.__init__(o)
    out["vboc1"] = (o.model.f_expl_expr, list(o.model.x), list(o.model.u))
    m = load("VBOC/doublependulum_class_vboc.py", "ref_p2_vboc")
    o = m.OCPdoublependulum()
    out["vboc2"] = (o.model.f_expl_expr, list(o.model.x), list(o.model.u))
    m = load("VBOC/triplependulum_class_vboc.py", "ref_p3_vboc")
    o = m.OCPtriplependulum()
    out["vboc3"] = (o.model.f_expl_expr, list(o.model.x), list(o.model.u))
    for n, rel, cls in ((1, "AL/pendulum_class_al.py", "OCPpendulum"),
                        (2, "AL/doublependulum_class_al.py", "OCPdoublependulum"),
                        (3, "AL/triplependulum_class_al.py", "OCPtriplependulum")):
        m = load(rel, f"ref_p{n}_al")
        o = getattr(m, cls)()
        out[f"al{n}"] = (o.model.f_expl_expr, list(o.model.x), list(o.model.u))
    return out
