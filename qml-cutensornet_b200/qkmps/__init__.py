"""ctypes binding of libqkmps.so (C ABI: include/qkmps.h) -- the B200 quantum-kernel MPS engine.

There is no CPU fallback: importing works anywhere (so CPU-only tests can check the ABI), but every
compute call needs the CUDA library and an sm_100 device and raises ``QkError`` otherwise.
"""

from __future__ import annotations

import ctypes
import os
import pathlib
import subprocess

import numpy as np

_HERE = pathlib.Path(__file__).resolve().parent
LIB_PATH = pathlib.Path(os.environ.get("QKMPS_LIB", _HERE / "libqkmps.so"))   # QKMPS_LIB: tuning builds only
CSRC = _HERE.parent / "csrc"

QK_TRUNC_ITENSORS = 0
QK_TRUNC_PYTKET = 1
QK_FLAG_CAP_HIT = 1
QK_FLAG_NO_CONVERGE = 2
QK_ERR_LIMIT = -3
QK_PLAN_LITERAL_ORDER = 1
QK_PLAN_EARLY_EXIT = 2
QK_PLAN_NO_FUSION = 4
QK_PLAN_PARALLEL = 8
CHI_LIMIT = 512         # stage-1 kernels: shared-memory-resident up to 32, large-matrix (cluster) kernel above
QK_PLAN_BIG = 16
DMMA_D_LIMIT = 16       # register-resident tensor-core overlap kernel

GATE_KIND = {"H": 0, "Rz": 1, "Rx": 2, "XXPhase": 3, "ZZPhase": 4, "SWAP": 5}


class QkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libqkmps error {code}: {msg}")
        self.code = code


class QkGate(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("q0", ctypes.c_int32), ("q1", ctypes.c_int32),
                ("fa", ctypes.c_int32), ("fb", ctypes.c_int32), ("coeff", ctypes.c_double)]


class QkPlanInfo(ctypes.Structure):
    _fields_ = [("n_qubits", ctypes.c_int32), ("n_gates", ctypes.c_int32), ("n_ops", ctypes.c_int32),
                ("n_ops_2q", ctypes.c_int32), ("n_ops_1q", ctypes.c_int32), ("n_moves", ctypes.c_int32),
                ("chi_cap", ctypes.c_int32), ("threads", ctypes.c_int32), ("trunc_mode", ctypes.c_int32),
                ("smem_bytes", ctypes.c_int32), ("state_stride", ctypes.c_int64), ("trunc_error", ctypes.c_double)]


class QkOpView(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("site", ctypes.c_int32), ("fa", ctypes.c_int32),
                ("fb", ctypes.c_int32), ("dir", ctypes.c_int32), ("coeff", ctypes.c_double)]


EXPORTS = [
    "qk_version", "qk_last_error", "qk_device_count",
    "qk_plan_create_gates", "qk_plan_create_ansatz", "qk_plan_info", "qk_plan_ops", "qk_plan_destroy",
    "qk_simulate", "qk_simulate_dev", "qk_simulate_trace", "qk_batch_sim_ms", "qk_batch_size", "qk_batch_info", "qk_batch_export",
    "qk_batch_import", "qk_batch_max_chi", "qk_batch_destroy", "qk_frag_stride", "qk_batch_pack",
    "qk_batch_pack_scatter",
    "qk_gram_frags", "qk_batch_store", "qk_gram_lane", "qk_gram_store", "qk_gram_host", "qk_dmma_peak", "qk_pipe_mix", "qk_gram_big", "qk_batch_repack", "qk_simulate_async", "qk_batch_unit_seconds", "qk_batch_pack_async",
    "qk_gram_set_tile_clocks", "qk_gram_tile_clocks_used", "qk_batch_flags", "qk_batch_release_store",
]

_lib = None


def build(force: bool = False) -> pathlib.Path:
    """Compile libqkmps.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    srcs = [CSRC / f for f in ("qk_api.cu", "qk_sim.cu", "qk_gram.cu", "qk_plan.cpp", "qk_types.h",
                               "qk_sim_core.h", "qk_sim_big.h", "qk_plan.h", "qk_kernels.cuh")]
    srcs.append(_HERE.parent.parent / "include" / "qkmps.h")
    stale = force or not LIB_PATH.exists() or any(s.stat().st_mtime > LIB_PATH.stat().st_mtime for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", str(CSRC)], stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib() -> ctypes.CDLL:
    """Load libqkmps.so; fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise QkError(-2, f"{LIB_PATH} is missing: build it with `make -C {CSRC}` "
                          "(there is no CPU fallback for the quantum-kernel path)")
    L = ctypes.CDLL(str(LIB_PATH))
    L.qk_last_error.restype = ctypes.c_char_p
    L.qk_version.restype = ctypes.c_int
    for name in EXPORTS:
        getattr(L, name)  # AttributeError if the symbol is not exported
    _lib = L
    return L


def _check(rc: int):
    if rc < 0:
        raise QkError(rc, lib().qk_last_error().decode())
    return rc


def _p(arr, ctype):
    return arr.ctypes.data_as(ctypes.POINTER(ctype))


def gates_to_c(gates):
    """[(name, qubits, param)] with param None | ("lin", i, c) | ("prod", a, b, c) | ("const", alpha)."""
    arr = (QkGate * max(len(gates), 1))()
    for i, (name, qubits, param) in enumerate(gates):
        if name not in GATE_KIND:
            raise RuntimeError(f"Unrecognised {name}.")
        g = arr[i]
        g.kind = GATE_KIND[name]
        g.q0 = int(qubits[0])
        g.q1 = int(qubits[1]) if len(qubits) > 1 else -1
        g.fa, g.fb, g.coeff = -1, -1, 0.0
        if param is not None:
            if param[0] == "lin":
                g.fa, g.coeff = int(param[1]), float(param[2])
            elif param[0] == "prod":
                g.fa, g.fb, g.coeff = int(param[1]), int(param[2]), float(param[3])
            elif param[0] == "const":
                g.coeff = float(param[1])
            else:
                raise RuntimeError(f"bad gate parameter {param!r}")
    return arr


class Plan:
    """Compiled static op schedule of one ansatz (qk_plan)."""

    def __init__(self, n_qubits, gates, trunc_mode, trunc_error, chi_cap, flags=None):
        self._h = ctypes.c_void_p()
        carr = gates_to_c(gates)
        if flags is None:
            flags = 0
        if os.environ.get("QK_SCHEDULE", "") == "literal":
            flags |= QK_PLAN_LITERAL_ORDER
        if os.environ.get("QK_SCHEDULE", "") == "nofuse":
            flags |= QK_PLAN_NO_FUSION
        if os.environ.get("QK_SCHEDULE", "") == "parallel":
            flags |= QK_PLAN_PARALLEL
        _check(lib().qk_plan_create_gates(int(n_qubits), carr, len(gates), int(trunc_mode),
                                          ctypes.c_double(trunc_error), int(chi_cap), int(flags),
                                          ctypes.byref(self._h)))
        self.n_qubits = int(n_qubits)
        self.chi_cap = int(chi_cap)

    @classmethod
    def from_ansatz(cls, n_qubits, reps, gamma, pairs, hadamard_init, trunc_mode, trunc_error, chi_cap, flags=0):
        self = cls.__new__(cls)
        self._h = ctypes.c_void_p()
        pr = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1, 2))
        _check(lib().qk_plan_create_ansatz(int(n_qubits), int(reps), ctypes.c_double(gamma), int(bool(hadamard_init)),
                                           _p(pr, ctypes.c_int32), int(pr.shape[0]), int(trunc_mode),
                                           ctypes.c_double(trunc_error), int(chi_cap), int(flags),
                                           ctypes.byref(self._h)))
        self.n_qubits = int(n_qubits)
        self.chi_cap = int(chi_cap)
        return self

    def info(self) -> QkPlanInfo:
        out = QkPlanInfo()
        _check(lib().qk_plan_info(self._h, ctypes.byref(out)))
        return out

    def ops(self):
        n = self.info().n_ops
        arr = (QkOpView * max(n, 1))()
        k = _check(lib().qk_plan_ops(self._h, arr, n))
        return [(o.kind, o.site, o.fa, o.fb, o.dir, o.coeff) for o in arr[:k]]

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.qk_plan_destroy(self._h)
            self._h = None


class Batch:
    """Device-resident batch of simulated MPS (qk_batch)."""

    def __init__(self, handle):
        self._h = handle
        n_states, nq = ctypes.c_int(), ctypes.c_int()
        _check(lib().qk_batch_size(self._h, ctypes.byref(n_states), ctypes.byref(nq)))
        self.N, self.n_qubits = n_states.value, nq.value

    def info(self):
        N, n = self.N, self.n_qubits
        chi = np.ones((N, n + 1), dtype=np.int32)
        fid = np.ones(N)
        tw = np.zeros(N)
        nbytes = np.zeros(N, dtype=np.int64)
        flags = np.zeros(N, dtype=np.int32)
        sweeps = np.zeros(N, dtype=np.int32)
        _check(lib().qk_batch_info(self._h, _p(chi, ctypes.c_int32), _p(fid, ctypes.c_double), _p(tw, ctypes.c_double),
                                   _p(nbytes, ctypes.c_int64), _p(flags, ctypes.c_int32), _p(sweeps, ctypes.c_int32)))
        return dict(chi=chi, fidelity=fid, trunc_weight=tw, nbytes=nbytes, flags=flags, sweeps=sweeps)

    def sim_ms(self) -> float:
        ms = ctypes.c_float()
        _check(lib().qk_batch_sim_ms(self._h, ctypes.byref(ms)))
        return ms.value

    def release_store(self, stream: int = 0):
        """Give the site tensors' memory back (stream-ordered); bond dimensions and statistics stay readable."""
        _check(lib().qk_batch_release_store(self._h, ctypes.c_void_p(stream)))

    def flags_or(self) -> int:
        out = ctypes.c_int32()
        _check(lib().qk_batch_flags(self._h, ctypes.byref(out)))
        return out.value

    def unit_seconds(self) -> np.ndarray:
        """Seconds every datapoint's circuit took inside the stage-1 kernel (per-unit timing)."""
        out = np.zeros(self.N)
        _check(lib().qk_batch_unit_seconds(self._h, _p(out, ctypes.c_double)))
        return out

    def pack_async(self, D, frag_ptr: int, first_index: int, stream: int = 0):
        D = np.ascontiguousarray(D, dtype=np.int32)
        _check(lib().qk_batch_pack_async(self._h, _p(D, ctypes.c_int32), ctypes.c_void_p(frag_ptr), int(first_index),
                                         ctypes.c_void_p(stream)))

    def max_chi(self) -> np.ndarray:
        out = np.ones(self.n_qubits + 1, dtype=np.int32)
        _check(lib().qk_batch_max_chi(self._h, _p(out, ctypes.c_int32)))
        return out

    def export(self, i: int, chi_row=None):
        """Site tensors of state i as a list of [chi_l, 2, chi_r] complex128 arrays."""
        if chi_row is None:
            chi_row = self.info()["chi"][i]
        total = int(sum(int(chi_row[s]) * 2 * int(chi_row[s + 1]) for s in range(self.n_qubits)))
        buf = np.zeros(total, dtype=np.complex128)
        _check(lib().qk_batch_export(self._h, int(i), buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(buf.nbytes)))
        out, o = [], 0
        for s in range(self.n_qubits):
            cl, cr = int(chi_row[s]), int(chi_row[s + 1])
            out.append(buf[o:o + cl * 2 * cr].reshape(cl, 2, cr))
            o += cl * 2 * cr
        return out

    def pack(self, D, frag_ptr: int, stream: int = 0):
        D = np.ascontiguousarray(D, dtype=np.int32)
        _check(lib().qk_batch_pack(self._h, _p(D, ctypes.c_int32), ctypes.c_void_p(frag_ptr), ctypes.c_void_p(stream)))

    def pack_scatter(self, D, frag_ptr: int, dst_index, stream: int = 0):
        """Pack state i at position ``dst_index[i]`` of the frag buffer (negative: skip)."""
        D = np.ascontiguousarray(D, dtype=np.int32)
        dst = np.ascontiguousarray(dst_index, dtype=np.int32)
        assert dst.shape == (self.N,)
        _check(lib().qk_batch_pack_scatter(self._h, _p(D, ctypes.c_int32), ctypes.c_void_p(frag_ptr),
                                           _p(dst, ctypes.c_int32), ctypes.c_void_p(stream)))

    def repack(self, plan, store_ptr: int, chi_ptr: int, dst_index=None, stream: int = 0):
        """Copy the states into a store laid out for ``plan`` (state i -> position dst_index[i], negative: skip)."""
        dst = None
        if dst_index is not None:
            dst = np.ascontiguousarray(dst_index, dtype=np.int32)
            assert dst.shape == (self.N,)
        _check(lib().qk_batch_repack(self._h, plan._h, ctypes.c_void_p(store_ptr), ctypes.c_void_p(chi_ptr),
                                     None if dst is None else _p(dst, ctypes.c_int32), ctypes.c_void_p(stream)))

    def store(self):
        """(device pointer of the unpadded store, c128 per state, device pointer of chi [N][n+1] int32)."""
        ptr, chi, stride = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int64()
        _check(lib().qk_batch_store(self._h, ctypes.byref(ptr), ctypes.byref(stride), ctypes.byref(chi)))
        return int(ptr.value or 0), int(stride.value), int(chi.value or 0)

    def gram_store(self, other=None) -> tuple[np.ndarray, float]:
        """CUDA-core cross-check kernel: K[y, x] with y over ``other`` (or self)."""
        rows = self.N if other is None else other.N
        K = np.zeros((rows, self.N))
        ms = ctypes.c_float()
        _check(lib().qk_gram_store(self._h, None if other is None else other._h, _p(K, ctypes.c_double),
                                   ctypes.c_int64(self.N), ctypes.byref(ms)))
        return K, ms.value

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.qk_batch_destroy(self._h)
            self._h = None


def device_count() -> int:
    c = ctypes.c_int()
    _check(lib().qk_device_count(ctypes.byref(c)))
    return c.value


def simulate(plan: Plan, X: np.ndarray, device: int = 0) -> Batch:
    X = np.ascontiguousarray(X, dtype=np.float64)
    h = ctypes.c_void_p()
    _check(lib().qk_simulate(plan._h, int(device), _p(X, ctypes.c_double), int(X.shape[0]), int(X.shape[1]),
                             ctypes.byref(h)))
    return Batch(h)


def simulate_dev(plan: Plan, x_ptr: int, N: int, ldx: int, device: int = 0, stream: int = 0) -> Batch:
    h = ctypes.c_void_p()
    _check(lib().qk_simulate_dev(plan._h, int(device), ctypes.c_void_p(stream), ctypes.c_void_p(x_ptr), int(N), int(ldx),
                                 ctypes.byref(h)))
    return Batch(h)


def simulate_async(plan: Plan, x_ptr: int, N: int, ldx: int, device: int = 0, stream: int = 0) -> Batch:
    """Queue stage 1 on ``stream`` without waiting for it (qk_simulate_async)."""
    h = ctypes.c_void_p()
    _check(lib().qk_simulate_async(plan._h, int(device), ctypes.c_void_p(stream), ctypes.c_void_p(x_ptr), int(N), int(ldx),
                                   ctypes.byref(h)))
    return Batch(h)


def simulate_trace(plan: Plan, x, device: int = 0):
    """Simulate one datapoint and return (Batch, [(op kind, site, MPS size in MiB after the op)]).

    With ``log=True``-style use (see ``main_track_mem.py``) print ``f"MPS size (MiB)={mib}"`` per 2-qubit op.
    """
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(1, -1))
    ops = plan.ops()
    tr = np.zeros(max(len(ops), 1))
    h = ctypes.c_void_p()
    _check(lib().qk_simulate_trace(plan._h, int(device), _p(x, ctypes.c_double), int(x.shape[1]), _p(tr, ctypes.c_double),
                                   int(len(tr)), ctypes.byref(h)))
    return Batch(h), [(o[0], o[1], float(b) / 2 ** 20) for o, b in zip(ops, tr)]


def import_batch(states, device: int = 0) -> Batch:
    """Upload MPS made elsewhere: ``states`` = list of lists of [chi_l, 2, chi_r] arrays (test hook)."""
    N, n = len(states), len(states[0])
    chi = np.ones((N, n + 1), dtype=np.int32)
    chunks = []
    for i, ts in enumerate(states):
        for s, t in enumerate(ts):
            chi[i, s], chi[i, s + 1] = t.shape[0], t.shape[2]
            chunks.append(np.ascontiguousarray(t, dtype=np.complex128).reshape(-1))
    buf = np.concatenate(chunks)
    h = ctypes.c_void_p()
    _check(lib().qk_batch_import(int(device), n, N, _p(chi, ctypes.c_int32), buf.ctypes.data_as(ctypes.c_void_p),
                                 ctypes.c_int64(buf.nbytes), ctypes.byref(h)))
    return Batch(h)


def pad_dims(max_chi) -> np.ndarray:
    return ((np.asarray(max_chi, dtype=np.int64) + 7) // 8 * 8).astype(np.int32)


def frag_stride(n_qubits: int, D) -> int:
    D = np.ascontiguousarray(D, dtype=np.int32)
    out = ctypes.c_int64()
    _check(lib().qk_frag_stride(int(n_qubits), _p(D, ctypes.c_int32), ctypes.byref(out)))
    return out.value


def gram_frags(device, n_qubits, Dx, fragx_ptr, Nx, Dy, fragy_ptr, Ny, tiles, symmetric, k_ptr, ldk, stream=0,
               wait=True, tile_clocks=None) -> float:
    """Tensor-core Gram kernel on packed fragments.  ``wait=False``: queued on ``stream`` only (returns 0.0).
    ``tile_clocks`` = (device pointer, capacity): per-CTA-tile clock64 ticks are written there."""
    Dx = np.ascontiguousarray(Dx, dtype=np.int32)
    Dy = Dx if Dy is None else np.ascontiguousarray(Dy, dtype=np.int32)
    tiles = np.ascontiguousarray(np.asarray(tiles, dtype=np.int32).reshape(-1, 4))
    ms = ctypes.c_float()
    lib().qk_gram_set_tile_clocks(ctypes.c_void_p(tile_clocks[0] if tile_clocks else 0),
                                  ctypes.c_int64(tile_clocks[1] if tile_clocks else 0))
    if not wait:
        _check(lib().qk_gram_frags(int(device), ctypes.c_void_p(stream), int(n_qubits), _p(Dx, ctypes.c_int32),
                                   ctypes.c_void_p(fragx_ptr), int(Nx), _p(Dy, ctypes.c_int32),
                                   ctypes.c_void_p(fragy_ptr or 0), int(Ny), _p(tiles, ctypes.c_int32),
                                   int(tiles.shape[0]), int(bool(symmetric)), ctypes.c_void_p(k_ptr), ctypes.c_int64(ldk),
                                   None))
        lib().qk_gram_set_tile_clocks(None, ctypes.c_int64(0))
        return 0.0
    _check(lib().qk_gram_frags(int(device), ctypes.c_void_p(stream), int(n_qubits), _p(Dx, ctypes.c_int32),
                               ctypes.c_void_p(fragx_ptr), int(Nx), _p(Dy, ctypes.c_int32),
                               ctypes.c_void_p(fragy_ptr or 0), int(Ny), _p(tiles, ctypes.c_int32),
                               int(tiles.shape[0]), int(bool(symmetric)), ctypes.c_void_p(k_ptr), ctypes.c_int64(ldk),
                               ctypes.byref(ms)))
    return ms.value


def gram_lane(plan: Plan, device, max_chi, storex_ptr, chix_ptr, Nx, storey_ptr, chiy_ptr, Ny, tiles, symmetric, k_ptr, ldk,
              stream=0) -> float:
    """Lane-per-pair overlap kernel on the unpadded stores (all bond dimensions <= max_chi <= 4)."""
    tiles = np.ascontiguousarray(np.asarray(tiles, dtype=np.int32).reshape(-1, 4))
    ms = ctypes.c_float()
    _check(lib().qk_gram_lane(plan._h, int(device), ctypes.c_void_p(stream), int(max_chi), ctypes.c_void_p(storex_ptr),
                              ctypes.c_void_p(chix_ptr), int(Nx), ctypes.c_void_p(storey_ptr or 0),
                              ctypes.c_void_p(chiy_ptr or 0), int(Ny), _p(tiles, ctypes.c_int32), int(tiles.shape[0]),
                              int(bool(symmetric)), ctypes.c_void_p(k_ptr), ctypes.c_int64(ldk), ctypes.byref(ms)))
    return ms.value


def gram_big(plan: Plan, device, dims_x, dims_y, storex_ptr, chix_ptr, Nx, storey_ptr, chiy_ptr, Ny, tiles, symmetric, k_ptr,
             ldk, stream=0) -> float:
    """Batched-GEMM overlap sweep on the unpadded stores (any bond dimension within the plan's caps)."""
    tiles = np.ascontiguousarray(np.asarray(tiles, dtype=np.int32).reshape(-1, 4))
    dx = np.ascontiguousarray(dims_x, dtype=np.int32)
    dy = dx if dims_y is None else np.ascontiguousarray(dims_y, dtype=np.int32)
    ms = ctypes.c_float()
    _check(lib().qk_gram_big(plan._h, int(device), ctypes.c_void_p(stream), _p(dx, ctypes.c_int32), _p(dy, ctypes.c_int32),
                             ctypes.c_void_p(storex_ptr), ctypes.c_void_p(chix_ptr), int(Nx),
                             ctypes.c_void_p(storey_ptr or 0), ctypes.c_void_p(chiy_ptr or 0), int(Ny),
                             _p(tiles, ctypes.c_int32), int(tiles.shape[0]), int(bool(symmetric)), ctypes.c_void_p(k_ptr),
                             ctypes.c_int64(ldk), ctypes.byref(ms)))
    return ms.value


def gram_host(plan: Plan, X, Y=None, device: int = 0) -> np.ndarray:
    """Whole path through the C ABI with host buffers (qk_gram_host)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    rows = X.shape[0]
    yp, ny = None, 0
    if Y is not None:
        Y = np.ascontiguousarray(Y, dtype=np.float64)
        rows, ny, yp = Y.shape[0], Y.shape[0], _p(Y, ctypes.c_double)
    K = np.zeros((rows, X.shape[0]))
    _check(lib().qk_gram_host(plan._h, int(device), _p(X, ctypes.c_double), int(X.shape[0]), yp, int(ny),
                              int(X.shape[1]), _p(K, ctypes.c_double), ctypes.c_int64(X.shape[0])))
    return K


def dmma_peak(device: int = 0, iters: int = 20000) -> float:
    out = ctypes.c_double()
    _check(lib().qk_dmma_peak(int(device), int(iters), ctypes.byref(out)))
    return out.value


def pipe_mix(device: int = 0, iters: int = 20000):
    """(ms DMMA warps alone, ms DFMA warps alone, ms both) -- do the FP64 tensor and FMA pipes overlap?"""
    out = (ctypes.c_float * 3)()
    _check(lib().qk_pipe_mix(int(device), int(iters), out))
    return tuple(float(v) for v in out)
