"""Drop-in mirrors of the reference's per-system OCP classes.

`vboc_b200/shim/VBOC/` and `vboc_b200/shim/AL/` hold modules with the reference's module and class names
(`triplependulum_class_vboc.OCPtriplependulumINIT`, ...); putting one of the two directories on `sys.path`
instead of the reference's makes `from triplependulum_class_vboc import OCPtriplependulumINIT,
SYMtriplependulumINIT` resolve to the GPU engine.  One call = a batch of one problem; the batched
entry point is `vboc_b200.engine.BatchSolver`.
"""
