"""Stand-in for ``mpi4py`` on hosts without MPI (put ``qml-cutensornet_b200/compat`` on ``PYTHONPATH``).

The reference's drivers do ``from mpi4py import MPI`` and use ``MPI.COMM_WORLD`` (``Get_rank`` / ``Get_size``),
``MPI.Wtime`` and, inside its own backends only, pickled send / recv / reduce (main.py:1,17,164-185;
gpu_backend/kernel_state_ansatz.py:346-352,416-419,428).  The B200 backends need none of the point-to-point calls:
ranks are ``torchrun`` processes (one per GPU) and the collectives run over NCCL (``qkmps.comm``).  This module gives
the drivers the names they import; with ``WORLD_SIZE`` > 1 in the environment ``COMM_WORLD`` joins torch.distributed
on first use.
"""
from . import MPI  # noqa: F401
