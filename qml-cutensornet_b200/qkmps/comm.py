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


def reduce_sum_to_root(comm, K):
    """Ranks fill disjoint tiles of K; summing assembles the matrix on rank 0 (reference gpu:428)."""
    _need_torch_comm(comm)
    comm.dist.reduce(K, dst=0, op=comm.dist.ReduceOp.SUM, group=comm.group)
    return K
