"""The drop-in claim, executed: the reference's OWN driver scripts (/root/reference/main.py and main_no_test.py,
byte for byte, via runpy) run against this repository's backends -- `from mpi4py import MPI` is served by
compat/mpi4py, `from {gpu,cpu}_backend.kernel_state_ansatz import KernelStateAnsatz, build_kernel_matrix` by the
package -- on a synthetic Elliptic-shaped CSV, and the kernel matrices they save are compared with the exact
statevector.

/root/reference exists only in the build container, which has no GPU, and the product has no CPU arithmetic path;
so in this (CPU) test the device step behind the backends (`qkmps.engine.build_gram`) is replaced by an
oracle-backed stand-in.  What is exercised for real is everything the reference's driver touches: the mpi4py
stand-in, both backend modules' entry points, argument handling, return shapes (root rank), the profiling JSON
files, and the driver's own SVC sweep on the returned matrices.  The same entry points are run on the device by
tests/test_gpu_parity.py::test_main_driver_end_to_end.
"""
import json
import os
import pathlib
import runpy
import sys

import numpy as np
import pytest

import oracle
from oracle.gram_ref import gram_from_mps, simulate_batch

REF = pathlib.Path("/root/reference")
ROOT = pathlib.Path(__file__).resolve().parent.parent
PKG = ROOT / "qml-cutensornet_b200"

pytestmark = pytest.mark.skipif(not (REF / "main.py").exists(), reason="reference tree not present (GPU box)")


def _fake_build_gram(n, r, g, emap, mode):
    def build_gram(comm, plan_factory, n_qubits, X, Y=None, chi_cap=16, device=None, return_device=False,
                   structural_cap=False, checkpoint=None):
        assert n_qubits == n
        plan_factory(chi_cap)                              # the schedule compiler (C++) runs for real
        xs = simulate_batch(n, r, g, emap, X, mode=mode)
        ys = None if Y is None else simulate_batch(n, r, g, emap, Y, mode=mode)
        K = gram_from_mps(xs, ys)

        def info(ms):
            chi = np.array([[1] + m.bond_dims() + [1] for m in ms], dtype=np.int32)
            return dict(chi=chi, fidelity=np.array([m.fidelity for m in ms]), trunc_weight=np.zeros(len(ms)),
                        nbytes=np.array([m.nbytes() for m in ms]), flags=np.zeros(len(ms), dtype=np.int32),
                        sweeps=np.zeros(len(ms), dtype=np.int32), seconds=np.full(len(ms), 1e-3))
        prof = dict(sim_ms_x=1.0, sim_ms_y=0.0 if ys is None else 1.0, chi_cap=chi_cap, info_x=info(xs),
                    info_y=None if ys is None else info(ys), gram_ms=1.0, exchange_s=0.0, pair_seconds=None,
                    gram_kernel="oracle stand-in", stage1_schedule="oracle stand-in", launches=0)
        return K, prof
    return build_gram


@pytest.mark.parametrize("script,backend", [("main.py", "GPU"), ("main.py", "CPU"), ("main_no_test.py", "GPU")])
def test_reference_driver_runs_unmodified(script, backend, tmp_path, monkeypatch):
    n, r, g, d, n_ill, n_lic, seed = 8, 2, 0.5, 1, 20, 20, 3
    monkeypatch.chdir(tmp_path)
    monkeypatch.syspath_prepend(str(PKG))
    monkeypatch.syspath_prepend(str(PKG / "compat"))
    for mod in [m for m in sys.modules if m == "mpi4py" or m.startswith("mpi4py.")]:
        monkeypatch.delitem(sys.modules, mod)
    import importlib
    mk = importlib.import_module("make_synthetic_dataset")
    (tmp_path / "datasets").mkdir()
    mk.make(200, 600, seed=1).to_csv(tmp_path / "datasets" / "elliptic_synth.csv", index=False)
    emap = oracle.entanglement_graph(n, d)
    modname = f"{backend.lower()}_backend.kernel_state_ansatz"
    mod = importlib.import_module(modname)
    monkeypatch.setattr(mod, "build_gram", _fake_build_gram(n, r, g, emap, "pytket" if backend == "GPU" else "itensors"))
    monkeypatch.setattr(sys, "argv", [script, backend, str(n), str(r), str(g), str(d), str(n_ill), str(n_lic), str(seed),
                                      "elliptic_synth.csv"])
    ns = runpy.run_path(str(REF / script), run_name="__main__")          # the reference's file, unmodified
    import mpi4py
    assert str(PKG / "compat") in mpi4py.__file__                        # the stand-in served the import
    tag = f"Nf{n}_r{r}_g{g}_p0.0_nn{d}_mslinear_Ntr{n_ill}_s{seed}_elliptic_synth"
    k_train = np.load(tmp_path / "kernels" / f"train_{tag}.npy")
    assert k_train.shape == (32, 32)
    assert np.abs(k_train - oracle.statevector_gram(n, r, g, emap, ns["reduced_train_features"])).max() < 1e-8
    prof = json.load(open(tmp_path / f"train_{tag}.json"))
    for key in ("lenX", "median_circ_sim", "q1_circ_sim", "q3_circ_sim", "median_product", "ave max chi x"):
        assert key in prof
    if script == "main.py":
        k_test = np.load(tmp_path / "kernels" / f"test_{tag}.npy")
        assert k_test.shape == (8, 32)
        assert np.abs(k_test - oracle.statevector_gram(n, r, g, emap, ns["reduced_train_features"],
                                                       ns["reduced_test_features"])).max() < 1e-8
        assert np.load(tmp_path / "data" / f"test_{tag}.npy").shape == (11, 5)      # the driver's own SVC sweep ran
        assert (tmp_path / f"test_{tag}.json").exists()


def test_mpi4py_stand_in_surface(monkeypatch):
    monkeypatch.syspath_prepend(str(PKG))
    monkeypatch.syspath_prepend(str(PKG / "compat"))
    for mod in [m for m in sys.modules if m == "mpi4py" or m.startswith("mpi4py.")]:
        monkeypatch.delitem(sys.modules, mod)
    for v in ("RANK", "WORLD_SIZE", "OMPI_COMM_WORLD_RANK", "PMI_RANK", "SLURM_PROCID", "SLURM_NTASKS"):
        monkeypatch.delenv(v, raising=False)
    from mpi4py import MPI
    c = MPI.COMM_WORLD
    assert (c.Get_rank(), c.Get_size()) == (0, 1)
    assert c.bcast({"a": 1}) == {"a": 1} and c.reduce(np.ones(2)).tolist() == [1.0, 1.0]
    t0 = MPI.Wtime()
    assert MPI.Wtime() >= t0
    monkeypatch.setenv("RANK", "3")
    monkeypatch.setenv("WORLD_SIZE", "8")
    assert (c.Get_rank(), c.Get_size()) == (3, 8)
    with pytest.raises(NotImplementedError):
        c.send(None, dest=1)
