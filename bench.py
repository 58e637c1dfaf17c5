#!/usr/bin/env python
"""bench.py -- converged OCP solves / s on the reference's headline workload (BASELINE.json):
3-DOF VBOC (`VBOC/triplependulum_vboc.py:33-103, 110`): one SQP solve per sampled problem, N = 100.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA engine through the C-ABI)
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm: the oracle port on all host cores

A step = one pass of the hot path over one batch of `--batch` synthetic problems per GPU (weak
scaling: every rank samples its own problems, no inter-GPU traffic during the solves; one all-gather
of the boundary states afterwards, as the drivers do before the NN fit).

`value`  : converged solves / s with the problem data resident in HBM (device time, CUDA events on the
           solver's stream, max over ranks).
`e2e`    : the same through `vboc_solve_batch` with HOST buffers: H2D of guesses/bounds and D2H of
           trajectories/stats inside the timed region.
`roofline`: FP64.  achieved = algorithmic flops of the launch (per stage 7348 / 3395 /
           856 flops per linearisation / IPM iteration / merit evaluation, times the per-problem
           counters the kernel exports) / kernel time; peak = DFMA peak measured live on the device.
           `traffic` = DRAM bytes per launch: the ncu-measured bytes per IPM iteration of this kernel
           (profiles/r1_traffic.json, one `ncu --set full` capture) times the IPM iterations of the launch;
           `hbm` restates the same launch against the measured copy bandwidth (MEASURED_PEAKS.json).
`pipeline`: the full `data_generation` of `--pipeline` problems per GPU -- extension loop, retries, sub-OCP
           chains, twin simulation (VBOC/triplependulum_vboc.py:19-370) -- as the per-problem state machine on the
           device (`vboc_datagen_run`): SURVEY 8(d)(ii), 8(f)1; at N = 1 with the same generators on the CPU arm
           beside it (`pipeline.cpu`).
`other_configs`: (N = 1) C2 (`doublependulum_testdata.py`: 1 024 test points on the device state machine), C3 (2-DOF
           `data_generation` of 1 024 problems on the device state machine) and C5 (one AL round
           of the 3-DOF system: 46 656 SQP_RTI labels + the entropy query over the 15^6-state resident pool on the
           tcgen05 MLP kernel with device top-B and removal), so that they are on the driver's record too.
`cpu_baseline`: the CPU arm on the host cores, on a bounded sample of the same workload.  kind "acados" when
           `acados_template` + the reference scripts are importable on the box (tools/acados_arm.py: the UNMODIFIED
           reference classes under Pool(os.cpu_count())); otherwise kind "port": the oracle restatement, compiled on
           the box with -O3 -march=native and compile-time dimensions (oracle/Makefile, `_native/`), OpenMP over
           problems -- "restated CPU baseline, not acados".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_DOF, N_STAGES = 3, 100
F_LIN, F_IPM, F_SIM = 7348.0, 3395.0, 856.0  # SURVEY.md 8(d), VBOC n = 3, per stage
METRIC = "converged OCP solves/sec, 3-DOF VBOC"
TRAFFIC_FILE = "r2_traffic.json" if os.path.exists(os.path.join(ROOT, "profiles", "r2_traffic.json")) else "r1_traffic.json"
UNIT = "OCP/s"


def algorithmic_flops(out):
    return float(N_STAGES * (F_LIN * out["sqp_iter"].sum() + F_IPM * out["qp_iter"].sum()
                             + F_SIM * out["ls_evals"].sum()))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self._stop_evt.is_set():
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                self.rows.append([c.strip() for c in o.stdout.strip().split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            if len(r) < 6:
                continue
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


WORKLOAD = "triplependulum_vboc single SQP solve per problem, N=100, nx=7, nu=3"


def shared_config(batch, world):
    """`config` of BOTH arms (the reference arm runs on our arm's config; what differs between the arms -- the
    bounded CPU sample, the stream overlap -- is reported outside `config`)."""
    return {"workload": WORKLOAD, "batch_per_gpu_per_step": batch, "parallelism": f"dp{world} (index-sharded problems)",
            "sampler": "vboc_b200.problems.sample_vboc (VBOC/triplependulum_vboc.py:33-103), Philox seeds 10000*(rank+1)+step"}


_CPU_KIND = None


def cpu_backend():
    """('acados', None) when the reference's own solver path can run here, else ('port', oracle module) with the
    natively compiled baseline build of the oracle."""
    global _CPU_KIND
    if _CPU_KIND is None:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import acados_arm
        ok, why = acados_arm.available()
        if ok:
            _CPU_KIND = ("acados", acados_arm, why)
        else:
            from oracle import oracle as orc
            try:
                orc.use_native_baseline_build()
                how = "oracle port, -O3 -march=native, compile-time dimensions"
            except Exception as e:  # no compiler on the box: the portable build
                orc.lib()
                how = f"oracle port, portable -O2 build (native build failed: {e})"
            _CPU_KIND = ("port", orc, how + "; acados probe: " + why)
    return _CPU_KIND


def cpu_sample(nprob, seed, threads):
    """Time the CPU arm on `nprob` problems of the workload with `threads` host threads / processes."""
    from vboc_b200 import problems as pr
    kind, mod, _ = cpu_backend()
    bp = pr.sample_vboc(N_DOF, nprob, seed=seed)
    if kind == "acados":
        r = mod.solve_batch(N_DOF, bp, processes=threads)
        return int((r["status"] == 0).sum()), r["wall_s"]
    t0 = time.perf_counter()
    r = mod.solve_batch(N_DOF, mod.FAMILY_VBOC, mod.MODE_SQP, bp, nthreads=threads)
    dt = time.perf_counter() - t0
    return int((r["status"] == 0).sum()), dt


def cpu_pipeline(nprob, seed):
    """The full `data_generation` (VBOC/triplependulum_vboc.py:19-370) of `nprob` problems on the host cores: the
    same generators as the GPU pipeline leg over the oracle backend, every round's solves as one OpenMP batch."""
    from vboc_b200 import drivers
    kind, mod, _ = cpu_backend()
    if kind != "port":
        return None

    class Backend:
        N_max = drivers.N_CAP

        def solve(self, bp, mode):
            return mod.solve_batch(N_DOF, mod.FAMILY_VBOC, mode, bp)

    def sim(n, X, U, T):
        return np.stack([mod.rk4(n, 1, x, u, T) for x, u in zip(X, U)])

    st = {}
    t0 = time.perf_counter()
    rows = drivers.data_generation_batch(N_DOF, nprob, seed=seed, backend=(Backend(), sim), stats=st)
    dt = time.perf_counter() - t0
    return {"problems": nprob, "rows": int(rows.shape[0]), "solves": st.get("solves", 0), "converged": st.get("converged", 0),
            "wall_s": dt, "converged_solves_per_s": st.get("converged", 0) / dt, "rounds": st.get("rounds")}


def run_reference(args):
    """CPU arm: the reference's arithmetic lives in acados/HPIPM, which cannot be installed here
    (DESIGN.md), so this times the oracle port with all host threads; a step = a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    world = int(os.environ.get("WORLD_SIZE", "1"))
    kind, _, how = cpu_backend()
    # bounded sample per step: >= 64 problems per core so that the heavy tail of the iteration counts (0.5 % of the
    # problems run all 1000 SQP iterations = seconds on one core) amortises, sized from a short calibration so that
    # the whole --steps/--warmup run stays within a few minutes
    nprob = args.cpu_sample if args.cpu_sample > 0 else 64 * cores
    c0, dt0 = cpu_sample(16 * cores, 899, cores)
    rate0 = max(16 * cores / dt0, 1e-9)
    budget = args.cpu_budget / max(args.steps + args.warmup, 1)
    if args.cpu_sample <= 0 and nprob / rate0 > budget:
        nprob = max(8 * cores, int(rate0 * budget))
    for i in range(args.warmup):
        cpu_sample(nprob, 900 + i, cores)
    conv, tot = 0, 0.0
    for i in range(args.steps):
        c, dt = cpu_sample(nprob, 10_000 + i, cores)  # rank 0's problems of the GPU arm, first nprob of each step
        conv += c
        tot += dt
    v = conv / tot
    sample = (f"{nprob} problems per step (= {nprob / cores:.0f} per core) x {args.steps} steps: the first {nprob} problems "
              f"of each of rank 0's GPU-arm steps; {how}")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": shared_config(args.batch, world),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if args.pipeline > 0 and kind == "port":
        line["pipeline"] = cpu_pipeline(min(args.pipeline, args.cpu_pipeline), 77)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32768, help="problems per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="problems per step of the CPU arm (0 = 64 per host core, shrunk to fit --cpu-budget)")
    ap.add_argument("--cpu-budget", type=float, default=240.0, help="seconds the whole CPU arm may take")
    ap.add_argument("--extras", type=int, default=1, help="1: also measure the other BASELINE.json configs (N = 1 only)")
    ap.add_argument("--cpu-pipeline", type=int, default=256, help="problems of the CPU data_generation pipeline leg")
    ap.add_argument("--pipeline", type=int, default=1024,
                    help="problems PER GPU of the data_generation pipeline leg (0 = skip); the reference's round is 1000")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from vboc_b200 import engine, problems as pr

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.batch

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # One solver handle + CUDA stream per step: the K steps of the timed region are launched back to
    # back on K streams, so the tail of one step (the few problems that need hundreds of SQP
    # iterations keep single warps busy for seconds) overlaps the other steps instead of idling the GPU.
    nh = max(args.steps, 1)
    streams = [torch.cuda.Stream(device=local) for _ in range(nh)]
    sols = []
    for st in streams:
        sv = engine.BatchSolver(N_DOF, "vboc", B, N_STAGES, device=local)
        sv.set_stream(st.cuda_stream)
        sols.append(sv)
    peak_tf = engine.fp64_peak_tflops(local)

    def batch_for(step):  # every rank / step samples its own problems
        return pr.sample_vboc(N_DOF, B, seed=10_000 * (rank + 1) + step)

    def run_resident(batches):
        """upload (untimed) -> timed: all kernels launched on their streams -> elapsed from the first
        launch to the last completion, CUDA events on the launching streams."""
        for sv, bp in zip(sols, batches):
            sv.upload(bp)
        barrier()
        start = torch.cuda.Event(enable_timing=True)
        ends = [torch.cuda.Event(enable_timing=True) for _ in batches]
        start.record(streams[0])
        for sv, st, ev in zip(sols, streams, ends):
            sv.solve_resident_async(0)
            ev.record(st)
        for sv in sols[:len(batches)]:
            sv.sync()
        barrier()
        ms = max(start.elapsed_time(ev) for ev in ends)
        outs = [sv.download() for sv in sols[:len(batches)]]
        return ms, outs

    # ---- device-resident throughput: `value`
    for i in range(0, args.warmup, nh):
        run_resident([batch_for(-1 - j) for j in range(i, min(i + nh, args.warmup))])
    batches = [batch_for(i) for i in range(args.steps)]
    sampler = ClockSampler(local)
    sampler.start()
    total_ms, outs = run_resident(batches)
    clocks = sampler.stop()
    conv = sum(int((o["status"] == 0).sum()) for o in outs)
    flops = sum(algorithmic_flops(o) for o in outs)
    t_dev = total_ms * 1e-3

    # ---- end-to-end through the C-ABI with host buffers: `e2e` (one host thread per step; the C call
    # releases the GIL, so uploads, kernels and downloads of different steps overlap)
    h2d = sum(a.nbytes for k, a in batches[0].items() if isinstance(a, np.ndarray)
              and k in ("N", "x_guess", "u_guess", "p", "lbx0", "ubx0", "lbx", "ubx", "lbxN", "ubxN", "lbu", "ubu", "C0"))
    d2h = B * ((N_STAGES + 1) * 7 + N_STAGES * 3) * 8 + B * 64
    results = [None] * args.steps

    def worker(i):
        results[i] = sols[i].solve(batches[i])

    barrier()
    t0 = time.perf_counter()
    threads = [threading.Thread(target=worker, args=(i,)) for i in range(args.steps)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    conv_e = sum(int((o["status"] == 0).sum()) for o in results)
    if world > 1:  # the one exchange step: boundary states of all ranks, before the NN fit
        rows = torch.from_numpy(np.ascontiguousarray(
            np.concatenate([o["x"][:, 0, :6] for o in results]))).cuda()
        gathered = [torch.empty_like(rows) for _ in range(world)]
        dist.all_gather(gathered, rows)
    barrier()
    t_e2e = time.perf_counter() - t0

    # ---- whole-job aggregates: max time over ranks, sum of converged solves
    t_dev_local = t_dev
    agg = torch.tensor([t_dev, t_e2e, float(conv), float(conv_e), flops], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = agg[:2].clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tot = agg[2:].clone()
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        t_dev, t_e2e = tmax.tolist()
        conv, conv_e, flops_all = tot.tolist()
    else:
        flops_all = flops
    for sv in sols:  # free the batch workspaces before the pipeline leg
        sv.close()
    pipeline = None
    if args.pipeline > 0:
        # the full data_generation of `--pipeline` problems PER GPU (weak scaling) as the per-problem
        # state machine on the device (vboc_datagen_run): host -> inputs, one launch, rows back; then the all-gather of
        # the rows the drivers do before the NN fit.  Timed host to host, max over ranks.
        from vboc_b200 import drivers
        pst = {}
        dgen = engine.DataGenerator(N_DOF, args.pipeline, device=local)
        drivers.data_generation_device(N_DOF, min(64, args.pipeline), seed=76, device=local, dgen=dgen)  # warm-up
        inp = drivers.dg_inputs(N_DOF, args.pipeline, 77, first=rank * args.pipeline)
        barrier()
        tp0 = time.perf_counter()
        rows, dst = dgen.run(inp)
        if world > 1:
            from vboc_b200 import distributed as vd
            rows = vd.all_gather_rows(rows)
        barrier()
        tp = time.perf_counter() - tp0
        kms = dgen.last_kernel_ms
        dgen.close()
        pagg = torch.tensor([tp, float(dst["solves"].sum()), float(dst["converged"].sum()), float(dst["sim_steps"].sum()),
                             float((dst["status"] != 1).sum())], dtype=torch.float64, device="cuda")
        if world > 1:
            tmx = pagg[:1].clone()
            dist.all_reduce(tmx, op=dist.ReduceOp.MAX)
            psum = pagg[1:].clone()
            dist.all_reduce(psum, op=dist.ReduceOp.SUM)
            tp, (p_solves, p_conv, p_sim, p_ok) = float(tmx.item()), psum.tolist()
        else:
            _, p_solves, p_conv, p_sim, p_ok = pagg.tolist()
        pipeline = {"workload": "triplependulum_vboc data_generation (extensions, retries, sub-OCP chains, twin simulation): "
                                "per-problem state machine on the device, one warp per problem (vboc_datagen_run)",
                    "problems_per_gpu": args.pipeline, "problems": args.pipeline * world, "problems_ok": int(p_ok),
                    "rows": int(rows.shape[0]), "solves": int(p_solves), "converged": int(p_conv), "sim_steps": int(p_sim),
                    "wall_s": tp, "kernel_ms_rank0": kms, "converged_solves_per_s": p_conv / tp,
                    # when the problems finished (rank 0, device clock): the wall time is the longest chain -- a problem
                    # whose solves run into the 1000-iteration limit restarts up to 10 times, ~10 s each on one warp
                    "t_done_p50_p90_p99_max_s": [round(float(v) * 1e-6, 2) for v in np.percentile(dst["t_done_us"], [50, 90, 99, 100])]}
        # when the problems finished (rank 0, device clock).  The wall time is the LONGEST CHAIN: a problem whose solves
        # run into the 1000-iteration limit restarts up to 10 times, ~10 s each on one warp -- the reference algorithm's
        # tail.  Throughput while the GPU is still busy: converged solves of the first 90 % of the problems / their time.
        order = np.argsort(dst["t_done_us"])
        k90 = max(1, int(0.9 * len(order)))
        t90 = float(dst["t_done_us"][order[k90 - 1]]) * 1e-6
        pipeline["converged_solves_per_s_first_90pct_rank0"] = float(dst["converged"][order[:k90]].sum()) / max(t90, 1e-9)
    # ---- the other configs of BASELINE.json, measured in the same run so that they are on the driver's record
    # (N = 1 only): C3 2-DOF data generation, C5 one AL round (labelling + pool query), A12 MLP
    other = None
    if world == 1 and args.extras:
        from vboc_b200 import drivers, nn as vnn
        from vboc_b200._lib import MODE_RTI
        from vboc_b200.shim.my_nn import NeuralNetCLS
        other = {}
        t0 = time.perf_counter()
        st3 = {}
        rows3 = drivers.data_generation_device(2, 1024, seed=5, device=local, stats=st3)
        other["C3_doublependulum_vboc_data_generation"] = {
            "problems": 1024, "rows": int(rows3.shape[0]), "converged": st3["converged"], "wall_s": time.perf_counter() - t0,
            "converged_solves_per_s": st3["converged"] / (time.perf_counter() - t0)}
        t0 = time.perf_counter()
        st2 = {}
        Xt2 = drivers.testing_device(2, 1024, seed=1, device=local, stats=st2)
        other["C2_doublependulum_testdata"] = {
            "problems": 1024, "test_points": int(Xt2.shape[0]), "solves": st2["solves"], "converged": st2["converged"],
            "wall_s": time.perf_counter() - t0, "converged_solves_per_s": st2["converged"] / (time.perf_counter() - t0)}
        nA, BA = 3, 6 ** 6
        bpa = pr.sample_al(nA, BA, seed=3)
        sa = engine.BatchSolver(nA, "al", BA, 100, device=local)
        sa.solve(bpa, MODE_RTI)
        t0 = time.perf_counter()
        oa = sa.solve(bpa, MODE_RTI)
        dta = time.perf_counter() - t0
        sa.close()
        torch.manual_seed(0)
        net = vnn.MLP.from_torch(NeuralNetCLS(6, 500, 2), device=local)
        Xp = np.random.default_rng(0).uniform(-1, 1, (15 ** 6, 6)).astype(np.float32) * np.array([0.8, 0.8, 0.8, 10.5, 10.5, 10.5], dtype=np.float32) \
            + np.array([np.pi, np.pi, np.pi, 0, 0, 0], dtype=np.float32)
        t0 = time.perf_counter()
        rp = vnn.ResidentPool(Xp, device=local)
        t_up = time.perf_counter() - t0
        rp.score(net, 3.14, 5.0)
        t0 = time.perf_counter()
        ms_score = rp.score(net, 3.14, 5.0)
        idxq, _, _ = rp.select(BA)
        rp.remove_selected()
        t_query = time.perf_counter() - t0
        rp.close()
        net.close()
        other["C5_triplependulum_al_round"] = {
            "labelled_states": BA, "labelling_wall_s": dta, "labels_per_s": BA / dta, "viable_fraction": float((oa["status"] == 0).mean()),
            "pool_states": int(Xp.shape[0]), "pool_upload_once_s": t_up, "pool_score_kernel_ms": ms_score,
            "pool_score_topB_remove_wall_s": t_query, "selected": int(len(idxq)),
            "mlp_algorithmic_tflops": 2.0 * Xp.shape[0] * (6 * 500 + 500 * 500 + 500 * 2) / (ms_score * 1e-3) / 1e12}
    if rank == 0:
        cores = os.cpu_count() or 1
        c_n = args.cpu_sample if args.cpu_sample > 0 else 32 * cores   # ~10-30 s of CPU work
        c_conv, c_dt = cpu_sample(c_n, 10_000, cores) if world == 1 else (0, 1.0)
        achieved_tf = flops / t_dev_local / 1e12  # rank 0's launches
        traffic = hbm = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", TRAFFIC_FILE)))
            ipm_iters = float(sum(int(o["qp_iter"].sum()) for o in outs))
            total_bytes = tj["dram_bytes_per_ipm_iteration"] * ipm_iters
            traffic = total_bytes / args.steps
            peak_gbs = None
            try:
                peak_gbs = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
            except Exception:
                pass
            ach = total_bytes / t_dev_local / 1e9
            hbm = {"achieved": ach, "peak": peak_gbs, "unit": "GB/s", "frac": ach / peak_gbs if peak_gbs else None,
                   "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy bandwidth)" if peak_gbs else "unavailable",
                   "bytes_per_ipm_iteration": tj["dram_bytes_per_ipm_iteration"],
                   "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum of one bounded launch "
                             f"(profiles/{TRAFFIC_FILE}) scaled by this launch's IPM iterations"}
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": conv / t_dev, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": shared_config(B, world),
            "run": {"l2": "inputs + per-warp workspaces (>700 MB) exceed the 126 MB L2; no flush needed",
                    "steps_overlap": "the K timed steps run on K CUDA streams (K solver handles); elapsed = first launch to last completion",
                    "converged_fraction": conv / (B * world * args.steps)},
            "e2e": {"value": conv_e / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": args.steps,
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf if peak_tf else None, "traffic": traffic,
                         "peak_source": "DFMA micro-benchmark run live by bench.py before the timed region "
                                        "(vboc_fp64_peak; MEASURED_PEAKS.json has no FP64 entry); nominal 148 SMs x 64 FMA/clk x 2 x "
                                        "1.965 GHz = 37.2",
                         "flops_per_launch": flops / args.steps, "hbm": hbm},
            "clocks": clocks,
        }
        if pipeline is not None:
            line["pipeline"] = pipeline
        if other is not None:
            line["other_configs"] = other
        if world == 1:
            kind, _, how = cpu_backend()
            line["cpu_baseline"] = {"value": c_conv / c_dt, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": f"the first {c_n} problems (= {c_n // cores} per core) of step 0 of this run; {how}"}
            if pipeline is not None and args.cpu_pipeline > 0:
                cp = cpu_pipeline(min(args.pipeline, args.cpu_pipeline), 77)
                if cp:
                    line["pipeline"]["cpu"] = cp
                    line["pipeline"]["vs_cpu"] = pipeline["converged_solves_per_s"] / cp["converged_solves_per_s"]
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
