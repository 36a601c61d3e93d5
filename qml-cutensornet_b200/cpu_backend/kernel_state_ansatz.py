"""Drop-in for the reference's ``cpu_backend/kernel_state_ansatz.py``: same entry points, arguments,
return value, side effects (``<info_file>.json``, per-rank checkpoint under ``tmp/``) and profiling keys.  The reference ran this backend
on ITensors.jl through ``KernelPkg.compute_tile``; here the same truncation semantics (ITensors
relative ``cutoff``, no renormalisation -- cpu:262 -> KernelPkg.jl:68) are served by the sm_100a
kernels in libqkmps.so.  There is no CPU arithmetic path in this package.

``circuit_for_data`` returns the gate-tuple list ``[(name, [qubits], [alpha])]`` exactly like the
reference (cpu:96-131), and an unknown gate raises ``RuntimeError`` (cpu:129).
"""

from __future__ import annotations

import json
import os
import sys
from statistics import mean, median
from typing import Optional

import numpy as np

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG_ROOT not in sys.path:
    sys.path.insert(0, _PKG_ROOT)

from qkmps import QK_PLAN_EARLY_EXIT, QK_PLAN_PARALLEL, QK_TRUNC_ITENSORS, Plan  # noqa: E402
from qkmps.ansatz import KernelStateAnsatzBase, expected_chi, structural_chi_bound  # noqa: E402
from qkmps.comm import Wtime  # noqa: E402
from qkmps.engine import Checkpoint, build_gram  # noqa: E402

_KNOWN = ("H", "Rx", "Rz", "XXPhase", "ZZPhase", "SWAP")


class KernelStateAnsatz(KernelStateAnsatzBase):
    def circuit_for_data(self, feature_values):
        """List of ``(name, qubits, params)`` tuples with the feature values substituted."""
        bound = self._bind(feature_values)
        out = []
        for name, qubits, param in bound.gates:
            if name not in _KNOWN:
                raise RuntimeError(f"Unrecognised {name}.")
            out.append((name, list(qubits), [] if param is None else [param[1]]))
        return out


def build_kernel_matrix(mpi_comm, ansatz, X, Y=None, info_file="info_file", truncation_error: float = 1e-16,
                        number_of_tiles: Optional[int] = None, chi: Optional[int] = None) -> np.ndarray:
    """Kernel matrix of dimensions ``len(Y) x len(X)`` (``len(X) x len(X)`` when ``Y`` is None).

    ``number_of_tiles`` (default ``4 * n_procs`` like the reference, cpu:179) sets the tile count reported in the
    profiling JSON and the granularity of the checkpoint: every rank cuts its rows of K into
    ``number_of_tiles // n_procs`` row groups and rewrites ``tmp/checkpoint_rank_<rank>_<info_file>.npz`` after each
    (reference: ``.npy`` after every tile, cpu:212-233,279-282); a restarted run skips the finished groups and the
    file is deleted on success (cpu:326).  Circuits are simulated once, not once per tile (KernelPkg.jl:81-99).
    """
    n_procs, rank, root = mpi_comm.Get_size(), mpi_comm.Get_rank(), 0
    lenX = len(X)
    lenY = lenX if Y is None else len(Y)
    number_of_tiles = number_of_tiles if number_of_tiles is not None else 4 * n_procs
    tile_side = max(int(np.floor(np.sqrt(lenX * lenY / number_of_tiles))), 1)
    n_tiles = int(np.ceil(lenX / tile_side)) * int(np.ceil(lenY / tile_side))

    n_qubits = ansatz.ansatz_circ.n_qubits
    gates = ansatz.ansatz_circ.get_commands()
    for name, _, _ in gates:
        if name not in _KNOWN:
            raise RuntimeError(f"Unrecognised {name}.")
    # compiled schedules are kept on the ansatz: the train and the test kernel of one run share them
    plans = ansatz.__dict__.setdefault("_qk_plans", {})

    def plan_factory(cap, early_exit=False, parallel=False):
        key = (QK_TRUNC_ITENSORS, float(truncation_error), cap, bool(early_exit), bool(parallel), os.environ.get("QK_SCHEDULE", ""))
        if key not in plans:
            plans[key] = Plan(n_qubits, gates, QK_TRUNC_ITENSORS, float(truncation_error), cap,
                              (QK_PLAN_EARLY_EXIT if early_exit else 0) | (QK_PLAN_PARALLEL if parallel else 0))
        return plans[key]

    if chi is not None:
        cap0 = int(chi)
    else:
        bound = max(1, structural_chi_bound(ansatz.num_qubits, ansatz.reps, ansatz.entanglement_map))
        dist = max([abs(a - b) for a, b in ansatz.entanglement_map] or [0])
        est = expected_chi(ansatz.gamma, bound, float(truncation_error), n_terms=ansatz.reps * dist)
        cap0 = next((c for c in (4, 8, 16) if est <= c), 16)
    start_time = Wtime()
    bound_s = max(1, structural_chi_bound(ansatz.num_qubits, ansatz.reps, ansatz.entanglement_map))
    ckpt = Checkpoint(os.path.join("tmp", f"checkpoint_rank_{rank}_{os.path.basename(str(info_file))}.npz"),
                      groups=max(1, number_of_tiles // max(n_procs, 1)))
    build_kernel_matrix.last_checkpoint = ckpt
    ckpt.abort_after = getattr(build_kernel_matrix, "_abort_after", None)      # tests only
    K, prof = build_gram(mpi_comm, plan_factory, n_qubits, np.asarray(X), None if Y is None else np.asarray(Y),
                         chi_cap=cap0, structural_cap=(cap0 >= bound_s), checkpoint=ckpt)

    if rank == root:
        ix, iy = prof["info_x"], prof["info_y"]
        # per-circuit times from the in-kernel clocks (rank 0's shard); per-product times from the per-tile clocks of the
        # tensor-core kernel where recorded, else the batch average (reference cpu:264-269,296-312)
        t_x = [float(t) for t in ix["seconds"]] or [0.0]
        t_y = [float(t) for t in iy["seconds"]] if iy is not None else []
        n_pairs = lenX * (lenX + 1) // 2 if Y is None else lenX * lenY
        pair_s = prof.get("pair_seconds")
        if pair_s is None or not len(pair_s):
            pair_s = np.array([prof["gram_ms"] * 1e-3 / max(n_pairs // n_procs, 1)])
        chi_x = [int(c.max()) for c in ix["chi"]] or [1]
        chi_y = [int(c.max()) for c in iy["chi"]] if iy is not None else chi_x
        profiling_dict = dict()
        profiling_dict["lenX"] = (lenX, "entries")
        profiling_dict["lenY"] = (None if Y is None else lenY, "entries")
        profiling_dict["n_tiles"] = (n_tiles, "tiles")
        profiling_dict["truncation_error"] = (truncation_error, "")
        profiling_dict["vdots_per_tile"] = (tile_side ** 2, "entries")
        profiling_dict["total_time"] = (Wtime() - start_time, "seconds")
        profiling_dict["median_tile_time"] = ((Wtime() - start_time) / max(n_tiles, 1), "seconds")
        profiling_dict["median_circ_sim"] = (median(t_x + t_y), "seconds")
        profiling_dict["q1_circ_sim"] = (float(np.percentile(t_x + t_y, 25)), "seconds")
        profiling_dict["q3_circ_sim"] = (float(np.percentile(t_x + t_y, 75)), "seconds")
        profiling_dict["median_product"] = (float(np.median(pair_s)), "seconds")
        profiling_dict["q1_product"] = (float(np.percentile(pair_s, 25)), "seconds")
        profiling_dict["q3_product"] = (float(np.percentile(pair_s, 75)), "seconds")
        profiling_dict["ave max chi x"] = (mean(chi_x), "chi x")
        profiling_dict["ave max chi y"] = (mean(chi_y or [1]), "chi y")
        with open(info_file + ".json", "w") as fp:
            json.dump(profiling_dict, fp, indent=4)
    build_kernel_matrix.last_profile = prof
    return K
