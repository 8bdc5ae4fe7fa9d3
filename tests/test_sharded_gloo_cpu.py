"""world_size = 2 `gloo` test of the fold-sharded runner (spatpca_b200/distributed.py) on CPU.

The runner's host logic -- fold placement, the per-stage all-gather of score rows, the
reference-order argmin, the hand-over of (index1, index2) to the next stage, the broadcast of the
full-data fit -- is exercised with a stand-in engine that serves per-fold rows computed by the CPU
oracle (the GPU engine cannot run here).  Every rank must end with exactly the single-process
oracle result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackedEngine:
    """Implements the engine interface used by `run_sharded`; records which folds it was asked for."""

    def __init__(self, case, want):
        import torch
        self.torch = torch
        self.case, self.want = case, want
        self.p, self.K = case["x"].shape[0], case["K"]
        self.asked = []
        self.bufs, self.present, self.prepared = {}, set(), set()

    # ---- Omega in column blocks (stand-ins: one double per column, tagged by the owner's column index) ----
    def omega_begin(self):
        self.lp = self.torch.full((self.p,), -1.0, dtype=self.torch.float64)
        self.x1 = self.torch.full((self.p,), -1.0, dtype=self.torch.float64)
        self.om = self.torch.full((self.p,), -1.0, dtype=self.torch.float64)
        self.asked.append(("omega_begin",))
        return True

    def omega_solve_cols(self, c0, c1):
        self.lp[c0:c1] = self.torch.arange(c0, c1, dtype=self.torch.float64)
        self.asked.append(("omega_solve", c0, c1))

    def omega_refine_cols(self, c0, c1):
        # two Newton-Schulz steps, one per call, False once none is left.  The residual / correction of a block needs
        # every column of the current iterate: the other ranks' blocks must have arrived before each step
        ar = self.torch.arange(self.p, dtype=self.torch.float64)
        step = sum(1 for a in self.asked if a[0] == "omega_refine")
        if step >= 2:
            return False
        if step == 0:
            assert bool((self.lp == ar).all())
        else:
            assert bool((self.x1 == 3.0 * ar).all())
        self.x1[c0:c1] = (3.0 if step == 0 else 5.0) * ar[c0:c1]
        self.asked.append(("omega_refine", c0, c1))
        return True

    def omega_product_cols(self, c0, c1):
        # the product needs every corrected column, i.e. the second exchange must have happened
        assert bool((self.x1 == 5.0 * self.torch.arange(self.p, dtype=self.torch.float64)).all())
        self.om[c0:c1] = 2.0 * self.torch.arange(c0, c1, dtype=self.torch.float64)
        self.asked.append(("omega_product", c0, c1))

    def omega_buffer(self, which, c0, c1):
        return {0: self.lp, 1: self.om, 2: self.x1}[which][c0:c1]

    def omega_sync(self):
        pass

    def omega_finish(self):
        assert bool((self.om == 2.0 * self.torch.arange(self.p, dtype=self.torch.float64)).all())
        self.asked.append(("omega_finish",))

    def prepare_starts(self, folds):
        # the SVD starts hide under the replicated first stage: asked for right after omega_begin(), before any solve
        assert not any(a[0] == "omega_solve" for a in self.asked)
        self.asked.append(("starts", tuple(folds)))

    def prepare(self):
        assert ("omega_finish",) in self.asked
        self.asked.append(("prepare",))

    # ---- stage-1 inverses as jobs (stand-ins: 8 doubles tagged with the job id) ----
    def can_share_inverses(self):
        return True

    def prepare_fold(self, k):
        self.prepared.add(k)

    def inverse_buffer(self, k, i):
        if (k, i) not in self.bufs:
            self.bufs[(k, i)] = self.torch.zeros(8, dtype=self.torch.float64)
        return self.bufs[(k, i)]

    def build_inverse(self, k, i):
        assert k in self.prepared
        self.inverse_buffer(k, i).fill_(1000.0 * k + i)
        self.present.add((k, i))
        self.asked.append(("inv", k, i))

    def mark_inverse(self, k, i):
        self.present.add((k, i))

    def sync(self):
        pass

    def stage1_fold(self, k):
        self.asked.append(("s1", k))
        n1 = len(self.case["tau1"])
        for i in range(1, n1):        # every inverse of this fold must be here, with the right contents
            assert (k, i) in self.present, (k, i)
            assert float(self.bufs[(k, i)][3]) == 1000.0 * k + i
        return self.want["aux"]["cv1"][k].copy()

    def stage1_full(self, index1):
        self.asked.append(("f1", index1))

    def stage2_fold(self, k, index1):
        self.asked.append(("s2", k, index1))
        return self.want["aux"]["cv2"][k].copy()

    def stage2_full(self, index1, index2):
        self.asked.append(("f2", index1, index2))

    def stage3_fold(self, k, index1, index2):
        self.asked.append(("s3", k, index1, index2))
        return self.want["aux"]["cv3"][k].copy()

    def get_eigenfn(self):
        return self.want["estimated_eigenfn"]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from oracle import spatpca_ref as ref
    from oracle.spatpca_ref import arma_colsum
    from spatpca_b200 import distributed as D
    from spatpca_b200.synth import make_case

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = make_case("grid3d_small")
    want = ref.spatpca_cv(case["x"], case["Y"], case["M"], case["K"], case["tau1"], case["tau2"], case["gamma"],
                          case["nk"], 100, 1e-4, case["l2"])

    def select_index(cv):                  # host-only twin of spb_select_index (no GPU library needed)
        sums = arma_colsum(np.asarray(cv))
        return int(np.argmin(sums)), sums / cv.shape[0]

    eng = OracleBackedEngine(case, want)
    coll = D.TorchCollective()
    D.prepare_sharded(eng, coll, D.my_folds(case["M"], coll))
    got = D.run_sharded(eng, select_index, coll, case["M"], len(case["tau1"]), len(case["tau2"]),
                        len(case["gamma"]), case["tau1"], case["tau2"], case["gamma"])
    ok = True
    for key in ("selected_tau1", "selected_tau2", "selected_gamma"):
        ok &= got[key] == want[key]
    for key in ("cv_score_tau1", "cv_score_tau2", "cv_score_gamma"):
        ok &= bool(np.array_equal(got[key], want[key]))
    ok &= bool(np.array_equal(got["estimated_eigenfn"], want["estimated_eigenfn"]))
    # each rank ran only its own folds, and exactly one rank ran the full-data fit
    my_folds = sorted({a[1] for a in eng.asked if a[0] == "s1"})
    ok &= my_folds == [k for k in range(case["M"]) if D.fold_owner(k, world) == rank]
    did_full = any(a[0] == "f1" for a in eng.asked)
    ok &= did_full == (rank == D.full_fit_owner(case["M"], world))
    ok &= ("starts", tuple(k for k in range(case["M"]) if D.fold_owner(k, world) == rank)) in eng.asked
    # Omega: this rank solved and multiplied exactly its own column block, then prepare() ran
    c0, c1 = D.omega_columns(eng.p, world)[rank]
    ok &= ("omega_solve", c0, c1) in eng.asked and ("omega_product", c0, c1) in eng.asked
    ok &= sum(1 for a in eng.asked if a == ("omega_refine", c0, c1)) == 2
    ok &= sum(1 for a in eng.asked if a[0] == "omega_solve") == 1 and ("prepare",) in eng.asked
    # the inverse jobs were spread as planned, and each rank built exactly its share
    jobs, owner = D.plan_inverse_jobs(case["M"], len(case["tau1"]), world)
    built = sorted((a[1], a[2]) for a in eng.asked if a[0] == "inv")
    ok &= built == sorted(j for j in jobs if owner[j] == rank)
    with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
        f.write("ok" if ok else "FAIL %r" % (got,))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_run_sharded_world2_gloo(tmp_path, world):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"rank{r}.txt").read() == "ok"


def test_inverse_job_plan():
    from spatpca_b200 import distributed as D
    for world in (1, 2, 3, 4, 8):
        jobs, owner = D.plan_inverse_jobs(5, 11, world)
        assert len(jobs) == 50 and set(owner) == set(jobs)
        loads = [sum(1 for j in jobs if owner[j] == r) for r in range(world)]
        assert max(loads) == -(-50 // world)                       # no rank above ceil(total / world)
        kept = sum(1 for (k, i) in jobs if owner[(k, i)] == D.fold_owner(k, world))
        assert kept >= min(50, sum(min(-(-50 // world), 10 * sum(1 for k in range(5) if D.fold_owner(k, world) == r))
                                   for r in range(world)))


def test_omega_column_split():
    from spatpca_b200 import distributed as D
    for p in (1, 15, 50, 900, 4096, 10000, 40000):
        for world in (1, 2, 3, 4, 8):
            cols = D.omega_columns(p, world)
            assert len(cols) == world and cols[0][0] == 0 and cols[-1][1] == p
            for (a, b), (c, d) in zip(cols, cols[1:]):
                assert b == c and a <= b and c <= d
            assert all(a % 16 == 0 for a, _ in cols if a < p)


def test_fold_placement():
    from spatpca_b200 import distributed as D
    for world in (1, 2, 4, 8):
        owners = [D.fold_owner(k, world) for k in range(5)]
        assert sorted(set(owners)) == list(range(min(world, 5)))
        fr = D.full_fit_owner(5, world)
        loads = [owners.count(r) for r in range(world)]
        assert loads[fr] == min(loads)
    assert D.full_fit_owner(5, 8) == 7 and D.full_fit_owner(5, 2) == 1 and D.full_fit_owner(5, 1) == 0
