"""Run under torchrun (one process per GPU): Gram matrix through the reference-facing entry point on
all ranks, compared on rank 0 with the exact statevector and the oracle.  Used by test_gpu_multi.py
and by hand:  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py
"""
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "qml-cutensornet_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))

import oracle  # noqa: E402
from gpu_backend.kernel_state_ansatz import KernelStateAnsatz, build_kernel_matrix  # noqa: E402
from qkmps.comm import init_from_env  # noqa: E402


def main():
    comm = init_from_env("nccl")
    rank, size = comm.Get_rank(), comm.Get_size()
    worst = 0.0
    kernels = set()
    for (n, r, g, d, nx, ny) in [(10, 2, 0.5, 1, 40, 13), (12, 2, 0.7, 2, 37, 9), (14, 2, 0.1, 2, 21, 21), (10, 3, 0.5, 4, 10, 6)]:
        emap = oracle.entanglement_graph(n, d)
        X = oracle.synthetic_features(nx, n, 0)
        Y = oracle.synthetic_features(ny, n, 1)
        ans = KernelStateAnsatz(n, r, g, emap)
        K = build_kernel_matrix(comm, ans, X, truncation_error=1e-16)
        kernels.add(build_kernel_matrix.last_profile["gram_kernel"])
        Kt = build_kernel_matrix(comm, ans, X, Y, truncation_error=1e-16)
        kernels.add(build_kernel_matrix.last_profile["gram_kernel"])
        if rank == 0:
            e1 = np.abs(K - oracle.statevector_gram(n, r, g, emap, X)).max()
            e2 = np.abs(Kt - oracle.statevector_gram(n, r, g, emap, X, Y)).max()
            assert K.shape == (nx, nx) and Kt.shape == (ny, nx)
            if not np.array_equal(K, K.T):
                bad = np.argwhere(K != K.T)
                print('ASYM', (n, r, g, d, nx, ny), len(bad), bad[:6].tolist(), np.abs(K - K.T).max(), flush=True)
            assert np.abs(K - K.T).max() < 1e-14
            worst = max(worst, e1, e2)
        else:
            assert K is None and Kt is None
    comm.Barrier()
    if rank == 0:
        assert worst < 1e-8, worst
        # both stage-2 paths were exercised across ranks: stores (chi <= 4) and packed fragments
        assert {"qk_gram_lane_kernel", "qk_gram_dmma_kernel"} <= kernels, kernels
        print(f"MULTI_GPU_OK ranks={size} max_err={worst:.3e}")


if __name__ == "__main__":
    main()
