"""Multi-GPU sharding of the GPAR fit: one process per GPU, tasks = (output i, restart r).

Training of output i conditions on the OBSERVED earlier outputs (examples/eeg.jl:53-92,212-281;
examples/GPAR_scaled_examples.jl:132-175 pass `[y1]`, `[y1, y2]`), so the per-output / per-restart
fits are mutually independent: they are partitioned over the ranks with NO data-path collective,
and torch.distributed (NCCL on GPUs, gloo in the CPU tests) only all-gathers the scalars
`(task id, minimum, minimizer)` at the end.  Only prediction is sequential over outputs:
the owner of output i broadcasts its posterior means to the later outputs
(GPAR_scaled_examples.jl:172 `[test_y1, y2_out]`).
"""
import numpy as np


def dist_info():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def shard_tasks(costs, world, rank=None):
    """Greedy longest-processing-time partition of tasks with the given relative costs (output 1 is
    a time-only LGSSM and much cheaper than the pseudo-point outputs; cost grows mildly with D).
    Deterministic: every rank computes the same assignment.  -> list of task indices per rank (or the
    list of `rank` when given)."""
    costs = np.asarray(costs, dtype=np.float64)
    order = np.argsort(-costs, kind="stable")
    load = np.zeros(world)
    assign = [[] for _ in range(world)]
    for t in order:
        r = int(np.argmin(load))          # ties -> lowest rank
        assign[r].append(int(t))
        load[r] += costs[t]
    for a in assign:
        a.sort()
    return assign if rank is None else assign[rank]


def gather_results(local, ntasks, nparam, device=None):
    """local: {task: (minimum, minimizer[nparam])} -> (minimum[ntasks], minimizer[ntasks, nparam]) on
    every rank.  One all_gather of a (max tasks per rank) x (2 + nparam) float64 tensor."""
    import torch
    import torch.distributed as dist
    rank, world = dist_info()
    if world == 1:
        vals = np.full(ntasks, np.nan); thetas = np.full((ntasks, nparam), np.nan)
        for t, (v, th) in local.items():
            vals[t] = v; thetas[t] = th
        return vals, thetas
    cap = (ntasks + world - 1) // world + ntasks % world + 1
    cap = ntasks          # LPT can be uneven; ntasks rows is tiny (48 B each)
    buf = torch.full((cap, 2 + nparam), float("nan"), dtype=torch.float64)
    for i, (t, (v, th)) in enumerate(sorted(local.items())):
        buf[i, 0] = t; buf[i, 1] = v; buf[i, 2:] = torch.as_tensor(np.asarray(th, dtype=np.float64))
    if device is not None:
        buf = buf.to(device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    vals = np.full(ntasks, np.nan); thetas = np.full((ntasks, nparam), np.nan)
    for o in out:
        o = o.cpu().numpy()
        for row in o:
            if row[0] == row[0]:
                vals[int(row[0])] = row[1]; thetas[int(row[0])] = row[2:]
    return vals, thetas


def broadcast_means(means, src, device=None):
    """Posterior means of one output passed down the GPAR chain (8 N* bytes)."""
    import torch
    import torch.distributed as dist
    rank, world = dist_info()
    if world == 1:
        return means
    t = torch.as_tensor(np.ascontiguousarray(means, dtype=np.float64))
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=src)
    return t.cpu().numpy()


def fit_tasks(tasks, costs, run_task, nparam, device=None, timing=None):
    """Runs `run_task(task) -> (minimum, minimizer)` for this rank's share of `tasks` and gathers
    everything.  -> (minimum[len(tasks)], minimizer[len(tasks), nparam]).  timing (dict, optional) receives
    "busy_seconds": this rank's own fitting time, before it waits for the others in the gather."""
    import time
    rank, world = dist_info()
    mine = shard_tasks(costs, world, rank)
    local = {}
    t0 = time.perf_counter()
    for ti in mine:
        local[ti] = run_task(tasks[ti])
    if timing is not None:
        timing["busy_seconds"] = time.perf_counter() - t0
        timing["tasks_this_rank"] = len(mine)
    return gather_results(local, len(tasks), nparam, device)


def best_per_output(tasks, minimum, minimizer):
    """tasks: list of (output, restart).  -> {output: (best minimum, its minimizer, restart)}"""
    best = {}
    for (o, r), v, th in zip(tasks, minimum, minimizer):
        if v == v and (o not in best or v < best[o][0]):
            best[o] = (float(v), np.array(th), r)
    return best


def row_slice_bounds(n_rows, world, align=1024):
    """Consecutive row slices of one sequence for `world` ranks: equal shares, boundaries rounded up to multiples of `align`
    (the library needs multiples of 4) -> world + 1 boundaries."""
    return [min(n_rows, (n_rows * i // world + align - 1) // align * align) for i in range(world)] + [n_rows]


def scaled_dtc_row_sharded(engine, k_time, k_out, theta, row_lo, grad=False, group=None, device=None):
    """ONE scaled-GPAR objective (compute_gpar_dtc_objective, src/gp/dtc.jl:83-128) whose rows are sharded over the RANKS of
    a torch.distributed group — one process per GPU, NCCL for the two (gradient: three) exchanges, every rank returns the
    result.  `engine` is this rank's Context holding the full (t, y), the pseudo-inputs and rows [row_lo, row_lo + N) of the
    inputs (its scaled_slice_* methods wrap the C ABI of the same name).  The carry of the M whitened columns crosses the
    slice boundaries in an all-gather of slice summaries; (beta'beta, beta'alpha) are summed by an all-reduce; the gradient
    adds an all-gather of tangent summaries and a 5-number all-reduce.  -> value or (value, gradient (5,))."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    kw = dict(dtype=torch.float64, device=device)

    def done(t):          # the library works on its own stream: a collective's result must be complete before it is handed over
        if t.is_cuda:
            torch.cuda.current_stream(t.device).synchronize()
        return t

    def all_gather(x):
        out = torch.empty(world * x.numel(), **kw)
        dist.all_gather(list(out.chunk(world)), x, group=group)
        return done(out)

    sc, stc = engine.scaled_slice_begin(k_time, k_out, theta, row_lo, grad)
    summary = torch.empty(sc, **kw)
    engine.scaled_slice_summary(summary)
    gathered = all_gather(summary)
    stats = torch.empty(stc, **kw)
    engine.scaled_slice_stats(gathered, rank, stats)
    dist.all_reduce(stats, group=group)
    done(stats)
    if not grad:
        return engine.scaled_slice_value(stats)
    summary2 = torch.empty(3 * sc, **kw)
    engine.scaled_slice_tangent_summary(stats, summary2)
    gathered2 = all_gather(summary2)
    s5 = torch.as_tensor(engine.scaled_slice_grad_partial(gathered2, rank), dtype=torch.float64).to(device if device is not None else "cpu")
    dist.all_reduce(s5, group=group)
    return engine.scaled_slice_grad_finish(done(s5).cpu().numpy())
