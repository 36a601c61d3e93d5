"""Process-level communication: replaces the reference's mpi4py pickled-object traffic
(gpu_backend/kernel_state_ansatz.py:346-352,416-419,428) with torch.distributed collectives on
device buffers (NCCL over NVLink on GPUs; gloo in the CPU-side tests).

``TorchComm`` exposes the slice of the mpi4py communicator API the reference's callers use
(Get_rank / Get_size / Barrier / bcast / reduce), so ``build_kernel_matrix(mpi_comm, ...)`` keeps its
signature.  ``Wtime`` mirrors ``MPI.Wtime``.
"""

from __future__ import annotations

import os
import time

import numpy as np


def Wtime() -> float:
    return time.perf_counter()


class TorchComm:
    def __init__(self, group=None):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised; call qkmps.comm.init_from_env() first")
        self.dist = dist
        self.group = group
        self.backend = dist.get_backend(group)

    def Get_rank(self):
        return self.dist.get_rank(self.group)

    def Get_size(self):
        return self.dist.get_world_size(self.group)

    def Barrier(self):
        self.dist.barrier(self.group)

    def _device(self):
        import torch
        return torch.device("cuda", torch.cuda.current_device()) if self.backend == "nccl" else torch.device("cpu")

    def bcast(self, obj, root=0):
        box = [obj]
        self.dist.broadcast_object_list(box, src=root, group=self.group)
        return box[0]

    def reduce(self, array, op=None, root=0):
        """Sum-reduce a numpy array to ``root`` (mpi4py lowercase ``reduce`` semantics: None elsewhere)."""
        import torch
        t = torch.from_numpy(np.ascontiguousarray(array)).to(self._device())
        self.dist.reduce(t, dst=root, op=self.dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy() if self.Get_rank() == root else None


def init_from_env(backend: str | None = None) -> TorchComm:
    """Initialise torch.distributed from torchrun's environment (RANK / WORLD_SIZE / MASTER_*)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")) % torch.cuda.device_count())
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group(backend=backend, rank=int(os.environ.get("RANK", "0")),
                                world_size=int(os.environ.get("WORLD_SIZE", "1")))
    return TorchComm()


def init_from_comm(comm) -> "TorchComm":
    """Bootstrap torch.distributed from an mpi4py-like communicator (``Get_rank`` / ``Get_size`` / ``bcast``), e.g.
    ``MPI.COMM_WORLD`` of the reference's drivers (main.py:17) launched with one process per GPU: rank 0 picks the
    rendezvous address, every rank joins a NCCL (GPU) or gloo (CPU tests) process group."""
    import torch
    import torch.distributed as dist
    if dist.is_initialized():
        return TorchComm()
    rank, size = comm.Get_rank(), comm.Get_size()
    addr = None
    if rank == 0:
        import socket
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        addr = (os.environ.get("MASTER_ADDR", "127.0.0.1"), int(os.environ.get("QK_MASTER_PORT", port)))
    if not hasattr(comm, "bcast"):
        raise TypeError("multi-rank runs need a communicator with bcast() (mpi4py) or a qkmps.comm.TorchComm; got %r"
                        % (type(comm),))
    addr = comm.bcast(addr, root=0)
    backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(rank % torch.cuda.device_count())
    dist.init_process_group(backend=backend, init_method=f"tcp://{addr[0]}:{addr[1]}", rank=rank, world_size=size)
    return TorchComm()


def _is_multi(comm) -> bool:
    return comm.Get_size() > 1


def _need_torch_comm(comm):
    if not isinstance(comm, TorchComm):
        raise TypeError("multi-rank runs need a qkmps.comm.TorchComm communicator (one process per GPU, "
                        "torch.distributed); got %r" % (type(comm),))


def allreduce_max_int(comm, value: int) -> int:
    if not _is_multi(comm):
        return int(value)
    return int(allreduce_max_array(comm, np.array([value], dtype=np.int32))[0])


def allreduce_max_array(comm, arr: np.ndarray) -> np.ndarray:
    if not _is_multi(comm):
        return np.asarray(arr)
    _need_torch_comm(comm)
    import torch
    t = torch.from_numpy(np.ascontiguousarray(arr)).to(comm._device())
    comm.dist.all_reduce(t, op=comm.dist.ReduceOp.MAX, group=comm.group)
    return t.cpu().numpy()


def allgather_bytes(comm, local):
    """All-gather equally sized uint8 device buffers into one contiguous buffer (rank order)."""
    _need_torch_comm(comm)
    import torch
    out = torch.empty(local.numel() * comm.Get_size(), dtype=local.dtype, device=local.device)
    comm.dist.all_gather_into_tensor(out, local, group=comm.group)
    return out


def allgather_into(comm, buf, offset: int, count: int):
    """In-place all-gather: every rank's ``buf[offset:offset+count]`` (offset = rank * count) lands in all ranks' ``buf``."""
    _need_torch_comm(comm)
    assert buf.numel() == count * comm.Get_size() and offset == comm.Get_rank() * count
    comm.dist.all_gather_into_tensor(buf, buf[offset:offset + count], group=comm.group)
    return buf


def gather_to_root(comm, local):
    """Gather equally sized device tensors to rank 0 (flat, rank order); ``None`` on the other ranks.  Replaces the
    reference's dense ``reduce(SUM)`` of the whole kernel matrix (gpu:428): every rank sends only its row panel."""
    _need_torch_comm(comm)
    import torch
    rank, size = comm.Get_rank(), comm.Get_size()
    flat = local.contiguous().view(-1)
    if rank == 0:
        out = torch.empty(size * flat.numel(), dtype=flat.dtype, device=flat.device)
        parts = list(out.split(flat.numel()))
        comm.dist.gather(flat, gather_list=parts, dst=0, group=comm.group)
        return out
    comm.dist.gather(flat, gather_list=None, dst=0, group=comm.group)
    return None


def reduce_sum_to_root(comm, K):
    """Ranks fill disjoint tiles of K; summing assembles the matrix on rank 0 (reference gpu:428)."""
    _need_torch_comm(comm)
    comm.dist.reduce(K, dst=0, op=comm.dist.ReduceOp.SUM, group=comm.group)
    return K
