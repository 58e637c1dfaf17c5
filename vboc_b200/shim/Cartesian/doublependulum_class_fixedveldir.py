"""Mirror of the reference's `VBOC/Cartesian constraints/doublependulum_class_fixedveldir.py:6-283`: the VBOC OCP of the
double pendulum with the Cartesian path constraint
    con_h_expr = (l1 sin th1 + l2 sin th2 - x_c)^2 + (l1 cos th1 + l2 cos th2 - y_c)^2,  lh = radius^2, uh = 1e6   (:150-158)
(the end effector stays outside a circle), solved by the CUDA engine (`vboc_set_cartesian`) instead of
acados.  Same class names, attributes (`radius`, `x_c`, `y_c`) and `OCP_solve(...)` signature as the reference; the driver
is `VBOC/Cartesian constraints/vboc_multiprocessing.py` (`testing`, `testing_test`: the data-generation workers of
`VBOC/doublependulum_vboc.py` around this class)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from vboc_b200.shim.VBOC import doublependulum_class_vboc as _base  # noqa: E402


class OCPdoublependulum(_base.OCPdoublependulum):
    def __init__(self):
        super().__init__()
        self.radius = self.l2 / 4
        self.x_c = 0
        self.y_c = -self.l1 - self.l2 / 2


class OCPdoublependulumINIT(_base.OCPdoublependulumINIT, OCPdoublependulum):
    def __init__(self):
        super().__init__()
        self.ocp_solver.cartesian = dict(xc=self.x_c, yc=self.y_c, radius=self.radius, uh=1e6)


SYMdoublependulumINIT = _base.SYMdoublependulumINIT
