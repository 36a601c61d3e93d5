"""``mpi4py.MPI`` names used by the reference's drivers, backed by the torchrun environment / torch.distributed."""
import os
import time

SUM = "SUM"
MAX = "MAX"
ANY_SOURCE = -1
ANY_TAG = -1


def Wtime() -> float:
    return time.perf_counter()


class _World:
    """``MPI.COMM_WORLD``: rank / size from the launcher's environment (torchrun: RANK / WORLD_SIZE; also the
    variables of mpirun / srun in case the processes were started by one of those)."""

    def __init__(self):
        self._tc = None

    @staticmethod
    def _env(names, default):
        for n in names:
            if n in os.environ:
                return int(os.environ[n])
        return default

    def Get_rank(self) -> int:
        return self._env(("RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", "SLURM_PROCID"), 0)

    def Get_size(self) -> int:
        return self._env(("WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", "SLURM_NTASKS"), 1)

    def as_torch_comm(self):
        """The torch.distributed communicator the B200 backends use (created on first use)."""
        if self._tc is None:
            os.environ.setdefault("RANK", str(self.Get_rank()))
            os.environ.setdefault("WORLD_SIZE", str(self.Get_size()))
            from qkmps.comm import init_from_env
            self._tc = init_from_env()
        return self._tc

    def Barrier(self):
        if self.Get_size() > 1:
            self.as_torch_comm().Barrier()

    def bcast(self, obj, root=0):
        return obj if self.Get_size() == 1 else self.as_torch_comm().bcast(obj, root=root)

    def reduce(self, array, op=SUM, root=0):
        return array if self.Get_size() == 1 else self.as_torch_comm().reduce(array, op=op, root=root)

    def _p2p(self, *a, **k):
        raise NotImplementedError("point-to-point pickled messages are not part of the B200 path: states are exchanged "
                                  "with one NCCL all-gather of packed device buffers (qkmps.engine)")

    send = recv = sendrecv = _p2p


COMM_WORLD = _World()
