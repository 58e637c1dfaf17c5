"""The active-learning loop of the AL drivers (AL/triplependulum_al.py:100-360, doublependulum_al.py, pendulum_al.py)
around the batched engine: SURVEY 8(f)2.

Per round the reference (i) scores the whole unlabeled pool with the classifier's entropy and takes the B most
uncertain states (:241-281), (ii) labels them with `testing_guess` -- one SQP_RTI solve each, started from the
trajectory the guess network predicts (:45-62, AL/triplependulum_class_al.py:171-201) -- under `Pool(30)`, (iii)
slides the training windows (`X_iter = X_iter[qi:] + new`, `X_traj = X_traj[qt:] + new`, :284-293) and (iv) retrains
both networks (:300-360).  Here (i) is the fused MLP + entropy kernel with a (global) top-B, (ii) one batched RTI
call; (iii) is this module; (iv) stays PyTorch (`fit_minibatch`, the reference's Adam loop).  Nothing here imports
the oracle: `label_fn` defaults to `drivers.al_label_batch` (GPU)."""
import numpy as np

from . import drivers


def fit_minibatch(model, optimizer, criterion, X, Y, mean, std, n_minibatch=512, loss_stop=0.1, it_max=1000,
                  beta=0.95, seed=0, normalize_targets=False):
    """The reference's training loop (AL/triplependulum_al.py:147-170 classifier, :176-201 guess network): random
    minibatches, Adam steps until the exponentially averaged loss drops below `loss_stop` or `it_max` steps."""
    import torch
    rng = np.random.default_rng(seed)
    dev = next(model.parameters()).device
    Xt = (torch.as_tensor(np.asarray(X), dtype=torch.float32, device=dev) - mean) / std
    Yt = torch.as_tensor(np.asarray(Y), dtype=torch.float32, device=dev)
    if normalize_targets:
        Yt = (Yt - mean) / std
    val, it = 1.0, 0
    while val > loss_stop and it <= it_max and len(Xt):
        ind = torch.as_tensor(rng.choice(len(Xt), size=min(n_minibatch, len(Xt)), replace=False), device=dev)
        optimizer.zero_grad(set_to_none=True)
        loss = criterion(model(Xt[ind]), Yt[ind])
        loss.backward()
        optimizer.step()
        val = beta * val + (1 - beta) * float(loss.item())
        it += 1
    return val, it


def predict_guess(model_guess, X, mean, std, N, nx):
    """`compute_problem_nnguess`'s initial guess (AL/triplependulum_class_al.py:180-192): stage 0 is the state itself,
    stages 1..N the de-normalised network output."""
    import torch
    dev = next(model_guess.parameters()).device
    with torch.no_grad():
        inp = (torch.as_tensor(np.asarray(X), dtype=torch.float32, device=dev) - mean) / std
        out = (model_guess(inp) * std + mean).cpu().numpy().astype(np.float64)
    xg = np.empty((len(X), N + 1, nx))
    xg[:, 0] = X
    xg[:, 1:] = out.reshape(len(X), N, nx)
    return xg


def _rows(X, labels, traj):
    """X_iter rows [x, one-hot label] (label 1 -> [0, 1], label 0 -> [1, 0]) and X_traj rows = the flattened
    (N+1) x nx trajectory of every viable sample, stage 0 being the sample itself -- the reference's layout
    (AL/triplependulum_al.py:24-43, :183-184: the guess network is trained with X_traj[i][:nx] as input and
    X_traj[i][nx:] as target).  Samples whose solve ended with any other status (label 2) produce NO row, as in
    the reference (`testing` returns nothing for them); their count is the third return value."""
    viable, unviable = labels == 1, labels == 0
    keep = viable | unviable
    onehot = np.where(viable[keep][:, None], [0.0, 1.0], [1.0, 0.0])
    it_rows = np.hstack([X[keep], onehot])
    tr_rows = traj[viable].reshape(int(viable.sum()), traj.shape[1] * traj.shape[2])
    return it_rows, tr_rows, int((~keep).sum())


def active_learning(n, pool, N_init, B, model, model_guess, mean, std, fit_cls, fit_guess, etp_stop=0.1,
                    max_rounds=100, N=100, Tf=1.0, device=0, label_fn=None, query_fn=None, history=None,
                    sharded=False, resident=False, guess_in_kernel=False):
    """Run the loop; returns (X_iter, X_traj, remaining pool).  `fit_cls(model, X_iter)` / `fit_guess(model_guess,
    X_traj)` retrain in place; `history` (list) receives one dict per round.

    sharded=True (under torch.distributed): `pool` is THIS rank's shard of the unlabeled pool and
    N_init / B are global numbers.  Every rank scores its shard, the query is the global top-B, each rank labels and
    removes the selected states that live in its shard, and the new rows are all-gathered so that all ranks hold
    identical training windows (and, with identical seeds, train identical networks).  `query_fn(model, pool, Bk)`
    must then return (local indices, entropies, global maximum of the selected entropies).

    resident=True: after the initial labelling the pool is uploaded ONCE to the device (`nn.ResidentPool`) and stays
    there; every round's scoring, top-B and removal run on the device (`drivers.al_query_resident`) and only the
    queried rows travel.  The host keeps the float64 states (the OCPs are solved from those, not from the float32
    copies the network sees) and the map from pool position to original row.

    guess_in_kernel=True: `compute_problem_nnguess`'s guess network is evaluated inside the solve kernel
    (`vboc_set_guess_network`) instead of by PyTorch + a (B, N+1, 2n) guess array through the host."""
    from . import distributed as vd
    from . import nn as vnn
    world = rank = 0
    if sharded:
        import torch.distributed as dist
        world, rank = dist.get_world_size(), dist.get_rank()

    def gather(it_rows, tr_rows, width_tr):
        if not sharded:
            return it_rows, tr_rows
        return vd.all_gather_rows(it_rows), vd.all_gather_rows(tr_rows.reshape(-1, width_tr))

    default_label = label_fn is None
    label_fn = label_fn or (lambda X, xg: drivers.al_label_batch(n, X, device=device, N=N, Tf=Tf, x_guess=xg))
    pool = np.asarray(pool, dtype=float)
    nx = 2 * n
    # initial labelling without a trained guess (testing(s0), :131-137)
    n0 = N_init
    if sharded:
        lo, hi = vd.shard_range(N_init, rank, world)
        n0 = min(hi - lo, len(pool))
    labels, traj = label_fn(pool[:n0], None)
    X_iter, X_traj, dropped = _rows(pool[:n0], labels, traj[:, :, :nx])
    X_iter, X_traj = gather(X_iter, X_traj, (N + 1) * nx)
    pool = pool[n0:]
    fit_cls(model, X_iter)
    if len(X_traj):
        fit_guess(model_guess, X_traj)
    k, etpmax = 0, 1.0
    rp = orig = None
    if resident:
        rp = vnn.ResidentPool(pool, device=device)
        orig = np.arange(len(pool), dtype=np.int64)   # pool position -> row of `pool` (the float64 states)
        alive = np.ones(len(pool), dtype=bool)

    def pool_left():
        n_loc = len(rp) if resident else len(pool)
        if not sharded:
            return n_loc
        return int(vd.all_gather_rows(np.array([[float(n_loc)]])).sum())

    while k < max_rounds:
        total = pool_left()
        if etpmax < etp_stop or total == 0:
            break
        Bk = min(B, total)
        if resident:
            net = vnn.MLP.from_torch(model, device=device)
            idx, _, etpmax = drivers.al_query_resident(rp, net, float(mean), float(std), Bk, sharded=sharded)
            net.close()
            rows = orig[idx]
            orig = np.delete(orig, idx)
            alive[rows] = False
            elems = pool[rows]
            k += 1
        elif query_fn is not None:
            q = query_fn(model, pool, Bk)
            idx, etp = q[0], q[1]
            etpmax = float(q[2]) if len(q) > 2 else float(np.max(np.asarray(etp)[np.asarray(idx, dtype=np.int64)]))
        else:
            net = vnn.MLP.from_torch(model, device=device)
            idx, etp, etpmax = drivers.al_query(net, pool, float(mean), float(std), Bk, sharded=sharded)
            net.close()
        if not resident:
            idx = np.asarray(idx, dtype=np.int64)
            k += 1
            elems = pool[idx]
            pool = np.delete(pool, idx, axis=0)
        use_guess = model_guess is not None and len(X_traj) and len(elems)
        in_kernel = guess_in_kernel and default_label and use_guess
        xg = predict_guess(model_guess, elems, mean, std, N, nx) if (use_guess and not in_kernel) else None
        if in_kernel:   # the guess network is evaluated inside the solve kernel (vboc_set_guess_network)
            labels, traj = drivers.al_label_batch(n, elems, device=device, N=N, Tf=Tf,
                                                  guess_net=(model_guess, float(mean), float(std)))
        elif len(elems):
            labels, traj = label_fn(elems, xg)
        else:
            labels, traj = np.zeros(0, dtype=np.int64), np.zeros((0, N + 1, nx))
        new_it, new_tr, drop = _rows(elems, labels, traj[:, :, :nx])
        dropped += drop
        new_it, new_tr = gather(new_it, new_tr, (N + 1) * nx)
        # sliding windows (:284-293): drop as many old rows as the batch size, append the new ones
        qt, qi = min(Bk, len(X_traj)), min(Bk, len(X_iter))
        X_traj = np.vstack([X_traj[qt:], new_tr]) if len(new_tr) or len(X_traj) else X_traj
        X_iter = np.vstack([X_iter[qi:], new_it])
        fit_cls(model, X_iter)
        if len(X_traj):
            fit_guess(model_guess, X_traj)
        if history is not None:
            history.append(dict(round=k, etpmax=etpmax, labelled=int(len(elems)), viable=int((labels == 1).sum()),
                                pool=int(len(rp) if resident else len(pool)), window=int(len(X_iter)),
                                traj_window=int(len(X_traj)), dropped_label2=int(dropped)))
    if resident:
        rp.close()
        pool = pool[alive]
    return X_iter, X_traj, pool
