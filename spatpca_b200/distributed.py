"""Fold-sharded spatpcaCV: one process per GPU, `torch.distributed` for the plumbing.

The reference's only parallel axis is the fold loop of each stage (the old parallelFor of
src/RcppSpatPCA.cpp:315, 385, 460) plus the independent full-data fit (:598-599, 661-666).
Folds are dealt round-robin to ranks; the full-data fit goes to the least-loaded rank.  The
three stages are separated by an argmin over fold-summed scores (:595-597, 657-659, 684-687),
so the only collective is an all-gather of M x |grid| doubles per stage, and one broadcast of
the p x K result.  There is no collective on the data path.

Selections are bitwise independent of the number of ranks: a fold's scores do not depend on
where the fold ran, and rows are summed in fold order k = 0..M-1 with the reference's
accumulation order (`spb_select_index`).
"""
from __future__ import annotations

import numpy as np


def fold_owner(k: int, world: int) -> int:
    return k % world


def full_fit_owner(M: int, world: int) -> int:
    """Rank with the fewest folds (ties: the highest such rank, which finishes its folds first)."""
    loads = [len([k for k in range(M) if fold_owner(k, world) == r]) for r in range(world)]
    best = min(loads)
    return max(r for r in range(world) if loads[r] == best)


def my_folds(M: int, coll) -> list:
    """The folds of the calling rank."""
    return [k for k in range(M) if fold_owner(k, coll.world) == coll.rank]


def plan_inverse_jobs(M: int, n_tau1: int, world: int):
    """Owner rank of every stage-1 inverse job (fold k, tau1 index i), i = 1..n_tau1-1.

    The job list is ordered by (i, k).  Every home rank first keeps up to ceil(total / world) jobs of
    its own folds (no transfer needed), the remaining jobs go greedily to the least-loaded rank.
    Deterministic: every rank computes the same plan."""
    jobs = [(k, i) for i in range(1, n_tau1) for k in range(M)]
    total = len(jobs)
    quota = -(-total // world) if world > 0 else total
    loads = [0] * world
    owner = {}
    for (k, i) in jobs:
        h = fold_owner(k, world)
        if loads[h] < quota:
            owner[(k, i)] = h
            loads[h] += 1
    for job in jobs:
        if job not in owner:
            r = min(range(world), key=lambda q: (loads[q], q))
            owner[job] = r
            loads[r] += 1
    return jobs, owner


class Collective:
    """Minimal collective interface; `TorchCollective` wraps torch.distributed (nccl or gloo)."""

    rank = 0
    world = 1

    def exchange(self, sends, recvs):
        """sends: [(dst_rank, buffer)], recvs: [(src_rank, buffer)] in one global order."""
        assert not sends and not recvs

    def broadcast_buffers(self, items):
        """items: [(src_rank, buffer)] in one global order; every rank ends with every buffer filled."""

    def all_gather_rows(self, rows: np.ndarray) -> np.ndarray:      # (n_local, width) -> (world, n_max, width)
        return rows[None]

    def broadcast(self, arr: np.ndarray, src: int) -> np.ndarray:
        return arr

    def barrier(self):
        pass


class TorchCollective(Collective):
    def __init__(self, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = device if device is not None else torch.device("cpu")

    def all_gather_rows(self, rows: np.ndarray) -> np.ndarray:
        t = self.torch.from_numpy(np.ascontiguousarray(rows, dtype=np.float64)).to(self.device)
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return np.stack([o.cpu().numpy() for o in out])

    def broadcast(self, arr: np.ndarray, src: int) -> np.ndarray:
        t = self.torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64)).to(self.device)
        self.dist.broadcast(t, src=src)
        return t.cpu().numpy()

    def barrier(self):
        self.dist.barrier()

    def _tensor(self, buf):
        """A torch view of an engine buffer: either already a tensor, or (device pointer, bytes)."""
        if isinstance(buf, self.torch.Tensor):
            return buf
        ptr, nbytes = buf

        class _DevArray:                      # zero-copy: the NCCL transfer reads / writes the engine's buffer
            __cuda_array_interface__ = {"shape": (nbytes // 8,), "typestr": "<f8", "data": (ptr, False),
                                        "version": 2}
        return self.torch.as_tensor(_DevArray(), device=self.device)

    def broadcast_buffers(self, items):
        """Column blocks of Lp / Omega: one NCCL broadcast per owner, straight from / into the engines' buffers."""
        keep = []
        works = []
        for src, buf in items:
            if not isinstance(buf, self.torch.Tensor) and buf[1] == 0:
                continue                      # nothing to exchange for this block (e.g. a replicated X0)
            t = self._tensor(buf)
            if t.numel() == 0:
                continue
            keep.append(t)
            works.append(self.dist.broadcast(t, src=src, async_op=True))
        for w in works:
            w.wait()
        if self.device.type == "cuda":
            self.torch.cuda.synchronize(self.device)

    def exchange(self, sends, recvs):
        """Point-to-point copies of whole p x p inverses between GPUs (NCCL send/recv over NVLink)."""
        ops = []
        keep = []
        for dst, buf in sends:
            t = self._tensor(buf)
            keep.append(t)
            ops.append(self.dist.P2POp(self.dist.isend, t, dst))
        for src, buf in recvs:
            t = self._tensor(buf)
            keep.append(t)
            ops.append(self.dist.P2POp(self.dist.irecv, t, src))
        if ops:
            for req in self.dist.batch_isend_irecv(ops):
                req.wait()
            if self.device.type == "cuda":
                self.torch.cuda.synchronize(self.device)


def omega_columns(p: int, world: int):
    """Column range [c0, c1) of Lp / Omega owned by each rank: an even split, boundaries on multiples of 16
    (128-byte aligned columns blocks for the GEMMs)."""
    step = -(-p // world)
    step = -(-step // 16) * 16
    return [(min(p, r * step), min(p, (r + 1) * step)) for r in range(world)]


def prepare_sharded(engine, coll: Collective, folds=None):
    """`engine.prepare()` with thinPlateSplineMatrix (:58-76) shared by the ranks: the LU of the bordered
    system is replicated, the solves, the residual / correction products of the refinement and the final
    product are split by columns, and the column blocks travel by NCCL broadcast (3 x 8 p^2 bytes in total).
    With one rank, or an engine without the column API, this is `engine.prepare()`.

    `folds`: this rank's folds.  The build is driven from the host (sync + exchange between its phases), so the SVD
    starts of the full data and of these folds are run right after omega_begin() has enqueued the replicated first
    stage (X0): they hide under it instead of following the build."""
    if coll.world > 1 and hasattr(engine, "omega_begin") and engine.omega_begin():
        if folds is not None and hasattr(engine, "prepare_starts"):
            engine.prepare_starts(list(folds))
        cols = omega_columns(engine.p, coll.world)
        c0, c1 = cols[coll.rank]
        engine.omega_solve_cols(c0, c1)
        engine.omega_sync()
        coll.broadcast_buffers([(r, engine.omega_buffer(0, a, b)) for r, (a, b) in enumerate(cols)])
        # refined build: one Newton-Schulz step (residual + correction of this rank's columns) per call, the updated
        # columns are exchanged after each; False once no step is left (or the build is not the refined one)
        while engine.omega_refine_cols(c0, c1):
            engine.omega_sync()
            coll.broadcast_buffers([(r, engine.omega_buffer(2, a, b)) for r, (a, b) in enumerate(cols)])
        engine.omega_product_cols(c0, c1)
        engine.omega_sync()
        coll.broadcast_buffers([(r, engine.omega_buffer(1, a, b)) for r, (a, b) in enumerate(cols)])
        engine.omega_finish()
    engine.prepare()


def _gather_stage(coll: Collective, local: dict, M: int, width: int) -> np.ndarray:
    """All-gather the per-fold score rows; returns the M x width matrix in fold order."""
    per_rank = (M + coll.world - 1) // coll.world
    buf = np.zeros((per_rank, width))
    mine = sorted(local)
    for slot, k in enumerate(mine):
        buf[slot, :] = local[k]
    allr = coll.all_gather_rows(buf)
    cv = np.zeros((M, width))
    for r in range(coll.world):
        ks = [k for k in range(M) if fold_owner(k, coll.world) == r]
        for slot, k in enumerate(ks):
            cv[k, :] = allr[r, slot, :]
    return cv


def _run_folds(engine, stage: str, folds, *idx) -> dict:
    """This rank's folds of one stage: concurrently (`stageN_folds`) when the engine offers it."""
    if not folds:
        return {}
    many = getattr(engine, stage + "_folds", None)
    if many is not None:
        return many(folds, *idx)
    one = getattr(engine, stage + "_fold")
    return {k: one(k, *idx) for k in folds}


def run_sharded(engine, select_index, coll: Collective, M: int, n_tau1: int, n_tau2: int, n_gamma: int,
                tau1, tau2, gamma, share_inverses: bool = True):
    """The stage sequence of spatpcaCV (:592-692) over a fold-sharded set of engines.

    `engine` is this rank's engine (already uploaded and prepared); `select_index(cv)` returns
    (argmin index, mean scores) like `spb_select_index`.  Returns the reference's result dict on
    every rank."""
    rank, world = coll.rank, coll.world
    mine = [k for k in range(M) if fold_owner(k, world) == rank]
    full_rank = full_fit_owner(M, world)

    # ---- stage 1 ----
    # whether the stage-1 systems are shared is decided by free device memory on each rank: every rank must take
    # the same branch (the exchange below is collective), so the ranks agree on the minimum first
    can_share = bool(share_inverses and world > 1 and n_tau1 > 1 and
                     getattr(engine, "can_share_inverses", lambda: False)())
    if world > 1:
        can_share = bool(np.min(coll.all_gather_rows(np.array([[1.0 if can_share else 0.0]]))) > 0.5)
    if can_share:
        # the (fold, tau1) inverses do not depend on the iterates: spread them over ALL ranks (also the
        # ones without a fold) and ship each to the fold's home rank; the chains then only iterate
        jobs, owner = plan_inverse_jobs(M, n_tau1, world)
        my_jobs = [j for j in jobs if owner[j] == rank]
        for k in sorted({k for (k, _) in my_jobs} | set(mine)):
            engine.prepare_fold(k)
        for (k, i) in my_jobs:
            engine.build_inverse(k, i)
        engine.sync()
        sends, recvs, arrived = [], [], []
        for (k, i) in jobs:
            src, dst = owner[(k, i)], fold_owner(k, world)
            if src == dst:
                continue
            if rank == src:
                sends.append((dst, engine.inverse_buffer(k, i)))
            if rank == dst:
                recvs.append((src, engine.inverse_buffer(k, i)))
                arrived.append((k, i))
        coll.exchange(sends, recvs)
        for (k, i) in arrived:
            engine.mark_inverse(k, i)
    rows = _run_folds(engine, "stage1", mine)
    index1 = 0
    if n_tau1 > 1:
        cv1 = _gather_stage(coll, rows, M, n_tau1)
        index1, score1 = select_index(cv1)
        sel_tau1 = float(tau1[index1])
    else:
        score1 = np.zeros(1)
        sel_tau1 = float(np.max(tau1))
    if rank == full_rank:
        engine.stage1_full(index1)

    # ---- stage 2 ----
    index2 = 0
    if n_tau2 > 1:
        rows = _run_folds(engine, "stage2", mine, index1)
        cv2 = _gather_stage(coll, rows, M, n_tau2)
        index2, score2 = select_index(cv2)
        sel_tau2 = float(tau2[index2])
    else:
        score2 = np.zeros(1)
        sel_tau2 = float(np.max(tau2))
    if rank == full_rank:
        engine.stage2_full(index1, index2)

    # ---- stage 3 ----
    rows = _run_folds(engine, "stage3", mine, index1, index2)
    cv3 = _gather_stage(coll, rows, M, n_gamma)
    index3, score3 = select_index(cv3)
    sel_gamma = float(gamma[index3]) if n_gamma > 1 else float(np.max(gamma))

    eig = engine.get_eigenfn() if rank == full_rank else np.zeros((engine.p, engine.K))
    eig = coll.broadcast(np.ascontiguousarray(eig), full_rank)
    return {
        "cv_score_tau1": score1, "cv_score_tau2": score2, "cv_score_gamma": score3,
        "estimated_eigenfn": np.asfortranarray(eig),
        "selected_tau1": sel_tau1, "selected_tau2": sel_tau2, "selected_gamma": sel_gamma,
        "index": (index1, index2, index3),
    }
