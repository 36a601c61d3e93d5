"""CPU tests of the multi-GPU host logic: sharding, Gram row-block partition and assembly, and the
torch.distributed plumbing (gloo, world_size 2) the NCCL path uses on the GPUs."""
import os
import socket

import numpy as np
import pytest

import oracle
from oracle.gram_ref import gram_from_mps, simulate_batch
from oracle.mps_ref import mps_inner


def test_shard_bounds_cover_everything():
    from qkmps.engine import shard_bounds
    for N in (0, 1, 7, 40, 1000, 1001):
        for size in (1, 2, 3, 4, 8):
            spans = [shard_bounds(N, size, r) for r in range(size)]
            assert spans[0][0] == 0 and spans[-1][1] == N
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            per = -(-N // size) if N else 0
            assert all(hi - lo <= per for lo, hi in spans)
            assert all(lo == min(r * per, N) for r, (lo, hi) in enumerate(spans))   # reference gpu:154 chunking


@pytest.mark.parametrize("symmetric", [True, False])
def test_panel_tiles_partition(symmetric):
    """Every (unordered, if symmetric) pair is computed by exactly one rank; rows stay inside the rank's own panel;
    the local block needs only the rank's own states; work is balanced across ranks."""
    from qkmps.engine import panel_tiles, shard_bounds
    shapes = [(1, 1), (13, 13), (40, 40), (100, 100), (1000, 1000)] + ([] if symmetric else [(40, 8), (100, 37), (7, 3)])
    for cols, rows in shapes:
        for size in (1, 2, 3, 4, 5, 8):
            cover = np.zeros((rows, cols), dtype=int)
            work = []
            for r in range(size):
                w = panel_tiles(cols, rows, symmetric, size, r)
                lo, hi = w["rows"]
                assert (lo, hi) == shard_bounds(rows, size, r)
                n = 0
                for kind in ("local", "remote"):
                    for r0, r1, c0, c1 in w[kind]:
                        assert lo <= r0 <= r1 <= hi and 0 <= c0 <= c1 <= cols
                        if kind == "local":
                            xlo, xhi = shard_bounds(cols, size, r)
                            assert xlo <= c0 and c1 <= xhi
                        for y in range(r0, r1):
                            xs = range(c0, min(c1, y + 1)) if (symmetric and kind == "local") else range(c0, c1)
                            for x in xs:
                                cover[y, x] += 1
                                n += 1
                work.append(n)
            if symmetric:
                both = cover + cover.T - np.diag(np.diag(cover))
                assert np.array_equal(both, np.ones((rows, cols), dtype=int)), (rows, size)
            else:
                assert np.array_equal(cover, np.ones((rows, cols), dtype=int))
            if rows >= 1000:
                assert max(work) <= 1.02 * (sum(work) / size) + 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import sys
    import pathlib
    root = pathlib.Path(__file__).resolve().parent.parent
    for p in (root, root / "qml-cutensornet_b200", root / "tests"):
        sys.path.insert(0, str(p))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    from qkmps.comm import (allgather_bytes, allgather_into, allreduce_max_array, allreduce_max_int, gather_to_root,
                            init_from_env)
    from qkmps.engine import panel_tiles, shard_bounds
    comm = init_from_env("gloo")
    assert comm.Get_rank() == rank and comm.Get_size() == world
    assert list(allreduce_max_array(comm, np.array([rank + 1, 5 - rank], dtype=np.int32))) == [world, 5]
    assert allreduce_max_int(comm, rank) == world - 1
    g = allgather_bytes(comm, torch.full((6,), rank, dtype=torch.uint8))
    assert g.tolist() == sum([[r] * 6 for r in range(world)], [])
    assert comm.bcast({"a": 1} if rank == 0 else None) == {"a": 1}

    # emulate the engine's flow with the oracle as the arithmetic: shard -> simulate -> gather -> tiles -> reduce
    n, r, gmm, d, N = 8, 2, 0.5, 1, 13
    emap = oracle.entanglement_graph(n, d)
    X = oracle.synthetic_features(N, n, 0)
    lo, hi = shard_bounds(N, world, rank)
    mine = simulate_batch(n, r, gmm, emap, X[lo:hi])
    gathered = [None] * world
    torch.distributed.all_gather_object(gathered, [m.tensors for m in mine])

    class _M:
        def __init__(self, t):
            self.tensors = t
    states = [_M(t) for part in gathered for t in part]
    assert len(states) == N
    # the engine's assembly: every rank fills its own row panel (global row index through an offset), the panels are
    # gathered to rank 0 and the symmetric matrix is completed with max(K, K^T)
    from qkmps.engine import _gather_panels
    work = panel_tiles(N, N, True, world, rank)
    per = -(-N // world)
    panel = torch.zeros((per, N), dtype=torch.float64)
    row0 = work["rows"][0]
    for kind in ("local", "remote"):
        for r0, r1, c0, c1 in work[kind]:
            for y in range(r0, r1):
                for x in (range(c0, min(c1, y + 1)) if kind == "local" else range(c0, c1)):
                    v = abs(mps_inner(states[y], states[x])) ** 2
                    panel[y - row0, x] = v
                    if kind == "local":
                        panel[x - row0, y] = v
    K = _gather_panels(comm, panel, N, N, per, True, torch)
    buf = torch.arange(world * 4, dtype=torch.uint8) * 0 + 255
    buf[rank * 4:(rank + 1) * 4] = rank
    allgather_into(comm, buf, rank * 4, 4)
    assert buf.tolist() == sum([[q] * 4 for q in range(world)], [])
    gr = gather_to_root(comm, torch.full((3,), float(rank)))
    assert (gr is None) == (rank != 0)
    if rank == 0:
        assert gr.tolist() == sum([[float(q)] * 3 for q in range(world)], [])
    red = comm.reduce(np.full((2, 2), float(rank + 1)))
    if rank == 0:
        Kref = gram_from_mps(simulate_batch(n, r, gmm, emap, X))
        assert np.abs(K.numpy() - Kref).max() < 1e-13
        assert np.array_equal(red, np.full((2, 2), float(sum(range(1, world + 1)))))
        out.put("ok")
    else:
        assert red is None
    comm.Barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_ranks(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert out.get(timeout=5) == "ok"


class _FileBcastComm:
    """mpi4py-like communicator for the test below: rank / size + a pickled bcast through a file (what the engine needs
    from a real ``MPI.COMM_WORLD`` to bootstrap torch.distributed)."""

    def __init__(self, rank, size, path):
        self.rank, self.size, self.path = rank, size, path

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.size

    def bcast(self, obj, root=0):
        import pickle
        import time
        if self.rank == root:
            with open(self.path + ".tmp", "wb") as f:
                pickle.dump(obj, f)
            os.replace(self.path + ".tmp", self.path)
            return obj
        for _ in range(600):
            if os.path.exists(self.path):
                with open(self.path, "rb") as f:
                    return pickle.load(f)
            time.sleep(0.05)
        raise TimeoutError("bcast")


def _worker_mpi_like(rank, world, path, out):
    import sys
    import pathlib
    root = pathlib.Path(__file__).resolve().parent.parent
    for p in (root, root / "qml-cutensornet_b200", root / "tests"):
        sys.path.insert(0, str(p))
    for v in ("RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        os.environ.pop(v, None)
    import torch
    from qkmps.comm import TorchComm, allreduce_max_int
    from qkmps.engine import _as_torch_comm
    comm = _as_torch_comm(_FileBcastComm(rank, world, path))       # what build_gram does with a foreign communicator
    assert isinstance(comm, TorchComm) and comm.Get_rank() == rank and comm.Get_size() == world
    assert allreduce_max_int(comm, rank + 10) == world + 9
    comm.Barrier()
    if rank == 0:
        out.put("ok")
    torch.distributed.destroy_process_group()


def test_mpi_like_communicator_bootstraps_torch_distributed(tmp_path):
    """A communicator that is not a TorchComm (mpi4py's COMM_WORLD with one process per GPU) is accepted at size > 1:
    the rendezvous address travels through its bcast."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    path = str(tmp_path / "bcast.pkl")
    procs = [ctx.Process(target=_worker_mpi_like, args=(r, 2, path, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert out.get(timeout=5) == "ok"
