"""`VBOC/pendulum_class_vboc.py` (1-DOF VBOC with a free `dt` state and a time term in the cost) is not supported
by the engine yet: constructing it says so instead of silently doing something else."""


class OCPpendulum:
    def __init__(self):
        raise NotImplementedError("1-DOF VBOC (free dt, VBOC/pendulum_class_vboc.py:52-105) is not supported by "
                                  "vboc_b200 yet; the 2/3-DOF classes and all AL classes are")
