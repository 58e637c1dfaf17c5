"""Mirror of the reference's `VBOC/Safe MPC/receiding_hard_constraints/doublependulum_class_fixedveldir.py:109-264`: the Safe-MPC OCP with the learned
margin `h(x) = out(x) (100 - safety_margin)/100 - max(|v|, 1e-3) >= 0` as a SOFT constraint at every stage
(`con_h_expr = con_h_expr_e`, `idxsh = idxsh_e = [0]`, slack penalties `Zl` set per stage at run time:
VBOC/Safe MPC/receiding_hard_constraints/2dof_sym.py:53-57 sets Zl = 1e12 at the receding stage and 10^((1 - r/N) 6) elsewhere).
Solved by the CUDA engine's MPC family with soft rows (`vboc_set_mpc_rows`).  Same class names and call
sequence as the reference; `ocp_solver.cost_set(i, "Zl", ...)` is served, `ocp_solver.get(i, "sl")` returns the slack."""
from vboc_b200.shim.SafeMPC import doublependulum_class_fixedveldir as _base


class OCPdoublependulumINIT(_base.OCPdoublependulumINIT):
    SOFT_ROWS = True


SYMdoublependulumINIT = _base.SYMdoublependulumINIT
