"""Test helper: drive the host emulation of the stage-1 core (tests/host_emu) through ctypes."""
from __future__ import annotations

import ctypes
import pathlib
import subprocess

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
EMU_DIR = ROOT / "tests" / "host_emu"
CSRC = ROOT / "qml-cutensornet_b200" / "csrc"

KIND = {"H": 0, "Rz": 1, "Rx": 2, "XXPhase": 3, "ZZPhase": 4, "SWAP": 5}


class QkGate(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("q0", ctypes.c_int32), ("q1", ctypes.c_int32),
                ("fa", ctypes.c_int32), ("fb", ctypes.c_int32), ("coeff", ctypes.c_double)]


def gates_to_c(gates):
    """oracle symbolic gate list -> ctypes array of qk_gate."""
    arr = (QkGate * len(gates))()
    for i, (name, qubits, param) in enumerate(gates):
        g = arr[i]
        g.kind = KIND[name]
        g.q0 = qubits[0]
        g.q1 = qubits[1] if len(qubits) > 1 else -1
        g.fa, g.fb, g.coeff = -1, -1, 0.0
        if param is not None:
            if param[0] == "lin":
                g.fa, g.coeff = param[1], param[2]
            elif param[0] == "prod":
                g.fa, g.fb, g.coeff = param[1], param[2], param[3]
            else:
                g.coeff = param[1]
    return arr


def build_emu() -> ctypes.CDLL:
    so = EMU_DIR / "libqk_emu.so"
    srcs = [EMU_DIR / "qk_emu.cpp", CSRC / "qk_plan.cpp"]
    deps = srcs + [CSRC / "qk_sim_core.h", CSRC / "qk_sim_big.h", CSRC / "qk_types.h", CSRC / "qk_plan.h"]
    if not so.exists() or any(d.stat().st_mtime > so.stat().st_mtime for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", str(so)] + [str(s) for s in srcs])
    lib = ctypes.CDLL(str(so))
    lib.qk_emu_simulate.restype = ctypes.c_longlong
    return lib


def emu_simulate(lib, n, gates, X, trunc_mode=0, trunc_error=1e-16, chi_cap=16, threads=0, flags=0):
    """Returns (list of per-state lists of site tensors, chi[N][n+1], stats[N][4], (n_ops, n_moves))."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    N = X.shape[0]
    carr = gates_to_c(gates)
    site_off = (ctypes.c_longlong * (n + 1))()
    n_ops, n_moves = ctypes.c_int(), ctypes.c_int()
    dptr = ctypes.POINTER(ctypes.c_double)
    stride = lib.qk_emu_simulate(n, carr, len(gates), trunc_mode, ctypes.c_double(trunc_error), chi_cap, flags, threads,
                                 X.ctypes.data_as(dptr), N, X.shape[1], None, None, site_off, None,
                                 ctypes.byref(n_ops), ctypes.byref(n_moves))
    if stride < 0:
        raise RuntimeError(f"qk_emu_simulate failed: {stride}")
    chi = np.zeros((N, n + 1), dtype=np.int32)
    store = np.zeros((N, stride), dtype=np.complex128)
    stats = np.zeros((N, 4), dtype=np.float64)
    rc = lib.qk_emu_simulate(n, carr, len(gates), trunc_mode, ctypes.c_double(trunc_error), chi_cap, flags, threads,
                             X.ctypes.data_as(dptr), N, X.shape[1],
                             chi.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                             store.ctypes.data_as(dptr), site_off, stats.ctypes.data_as(dptr),
                             ctypes.byref(n_ops), ctypes.byref(n_moves))
    if rc < 0:
        raise RuntimeError(f"qk_emu_simulate failed: {rc}")
    off = list(site_off)
    states = []
    for i in range(N):
        ts = []
        for s in range(n):
            cl, cr = chi[i, s], chi[i, s + 1]
            ts.append(store[i, off[s]:off[s] + cl * 2 * cr].reshape(cl, 2, cr).copy())
        states.append(ts)
    return states, chi, stats, (n_ops.value, n_moves.value)


class TensorsMPS:
    """Duck-typed stand-in for oracle.RefMPS (only .tensors is used by mps_inner)."""
    def __init__(self, tensors):
        self.tensors = tensors
        self.n = len(tensors)
