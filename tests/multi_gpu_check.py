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
    for (n, r, g, d, nx, ny) in [(10, 2, 0.5, 1, 40, 13), (12, 2, 0.7, 2, 37, 9), (14, 2, 0.1, 2, 21, 21), (10, 3, 0.5, 4, 10, 6),
                                   (14, 4, 1.0, 4, 11, 5)]:      # bond dimension up to 128: large-matrix stage 1 + batched GEMMs
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
    # ---- BASELINE-sized shards (round 2): 50 qubits, distance 2 (config 3 shape) with enough points that every
    # rank's shard takes its production path (tensor-core Gram kernel; B-form stage 1 for shards <= 150), a
    # rectangular train x test Gram of the same shape, and the published 165-qubit chi = 2 scaling shape.
    # Checked on rank 0 against the oracle on a sample of states (all pairs among them) and for exact symmetry.
    from oracle.gram_ref import gram_from_mps, simulate_batch
    big = 0.0
    for (n, r, g, d, nx, ny) in [(50, 2, 1.0, 2, 320, 0), (50, 2, 0.5, 2, 256, 96), (165, 2, 0.1, 1, 192, 0)]:
        emap = oracle.entanglement_graph(n, d)
        X = oracle.synthetic_features(nx, n, 0)
        Y = oracle.synthetic_features(ny, n, 1) if ny else None
        ans = KernelStateAnsatz(n, r, g, emap)
        K = build_kernel_matrix(comm, ans, X, Y, truncation_error=1e-16)
        prof = build_kernel_matrix.last_profile
        kernels.add(prof["gram_kernel"])
        if rank == 0:
            sx = np.unique(np.linspace(0, nx - 1, 12).round().astype(int))      # spread over every rank's shard
            refx = simulate_batch(n, r, g, emap, X[sx], mode="pytket")
            if ny:
                sy = np.unique(np.linspace(0, ny - 1, 6).round().astype(int))
                refy = simulate_batch(n, r, g, emap, Y[sy], mode="pytket")
                Kref, Ks = gram_from_mps(refx, refy), K[np.ix_(sy, sx)]
                assert K.shape == (ny, nx)
            else:
                Kref, Ks = gram_from_mps(refx), K[np.ix_(sx, sx)]
                assert K.shape == (nx, nx) and np.array_equal(K, K.T)
                assert np.abs(np.diag(K) - 1).max() < 1e-10
            e = float(np.abs(Ks - Kref).max())
            rel = float((np.abs(Ks - Kref) / np.maximum(Kref, 1e-300)).max())
            print(f"[multi] n={n} d={d} gamma={g} {nx}x{ny or nx}: max abs err {e:.2e}, max rel err {rel:.2e}, "
                  f"schedule: {prof['stage1_schedule']}, gram: {prof['gram_kernel']}", flush=True)
            big = max(big, e)
        else:
            assert K is None
    comm.Barrier()
    if rank == 0:
        assert big < 1e-8, big
        assert worst < 1e-8, worst
        # both stage-2 paths were exercised across ranks: stores (chi <= 4) and packed fragments
        assert {"qk_gram_lane_kernel", "qk_gram_dmma_kernel", "qk_big_gemm_kernel"} <= kernels, kernels
        print(f"MULTI_GPU_OK ranks={size} max_err={worst:.3e} max_err_baseline_shapes={big:.3e}")


if __name__ == "__main__":
    main()
