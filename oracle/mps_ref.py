"""Oracle: MPS simulation with the reference's truncation rules, and MPS overlap.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates what the reference delegates to third-party packages that are not in
/root/reference and cannot be installed offline:

* CPU backend, KernelPkg/src/KernelPkg.jl:45-72 ->
  ``ITensors.apply(gates, MPS(s, "0"); cutoff)`` (ITensors.jl 0.3.51,
  NDTensors 0.2.21; pinned in KernelPkg/Manifest.toml:288-292,461-465).
  Published algorithm (ITensors ``product(o, psi)`` / ``setindex!(psi, phi,
  r)`` / NDTensors ``truncate!``): for every gate, QR-orthogonalise the MPS
  to the gate's first site; 1-site gate: multiply it into the site tensor;
  adjacent 2-site gate: contract both sites with the gate, SVD, keep the
  spectrum selected by ``truncate!`` with a *relative* cutoff on the
  discarded squared weight, left site <- U, right site <- S V^dagger
  (orthogonality centre moves to the right site).  No renormalisation.
* GPU backend, gpu_backend/kernel_state_ansatz.py:141-144,221 ->
  pytket-cutensornet 0.6.0 ``simulate(.., MPSxGate, Config(truncation_fidelity
  = 1 - truncation_error))`` (README.md:32).  Published algorithm
  (``MPSxGate._apply_2q_gate``): canonicalise to the bond, SVD with
  ``abs_cutoff = value_of_zero = 1e-16``, keep the smallest number of
  singular values whose cumulative squared weight reaches the target
  fidelity, renormalise the kept ones, absorb S into the left factor and
  multiply ``mps.fidelity`` by the kept fraction.
* overlap, KernelPkg.jl:103-109 (``abs(inner(y, x))^2``) and gpu:380-383
  (``x_mps.vdot(y_mps)``): left-to-right transfer-matrix sweep.
"""

from __future__ import annotations

import numpy as np

from .ansatz import gate_matrix


# ----------------------------------------------------------------------------
# truncation predicates
# ----------------------------------------------------------------------------

def truncate_itensors(p: np.ndarray, cutoff: float, maxdim: int | None = None, mindim: int = 1) -> tuple[int, float]:
    """NDTensors ``truncate!`` with use_relative_cutoff=true, use_absolute_cutoff=false.

    ``p``: squared singular values sorted in decreasing order.  Returns
    (number kept, relative discarded weight).  Walk from the tail, discarding
    while ``discarded + p[n] <= cutoff * sum(p)`` and more than ``mindim``
    values remain.
    """
    p = np.asarray(p, dtype=np.float64)
    n = len(p)
    if n == 0 or p[0] <= 0.0:
        return 1, 0.0
    truncerr = 0.0
    if maxdim is not None:
        while n > maxdim:
            truncerr += p[n - 1]
            n -= 1
    scale = float(np.sum(p))
    if scale == 0.0:
        scale = 1.0
    while n > mindim and truncerr + p[n - 1] <= cutoff * scale:
        truncerr += p[n - 1]
        n -= 1
    return max(n, 1), truncerr / scale


def truncate_pytket(s: np.ndarray, truncation_fidelity: float, value_of_zero: float = 1e-16,
                    chi: int | None = None) -> tuple[int, float]:
    """pytket-cutensornet 0.6.0 ``_apply_2q_gate`` selection rule.

    ``s``: singular values, decreasing.  Values below ``value_of_zero`` are
    trimmed by the SVD itself (cuTensorNet ``abs_cutoff``); then singular
    values are taken from the largest until ``numer/denom`` reaches the
    target fidelity.  Returns (number kept, kept fraction of the weight).
    """
    s = np.asarray(s, dtype=np.float64)
    m = int(np.sum(s >= value_of_zero)) if value_of_zero > 0 else len(s)
    m = max(m, 1)
    s = s[:m]
    # total weight summed in the same (sequential, largest first) order as the running sum below, so that
    # numer == denom is reached as soon as the remaining values no longer change the sum -- a pairwise-summed
    # total differs from the running sum in the last bit and would keep rounding-noise values (sigma ~ 1e-16)
    denom = 0.0
    for v in s:
        denom += float(v) ** 2
    if denom == 0.0:
        return 1, 1.0
    if truncation_fidelity < 1.0:
        numer = 0.0
        k = 0
        while truncation_fidelity > numer / denom and k < m:
            numer += float(s[k] ** 2)
            k += 1
        k = max(k, 1)
    else:
        k, numer = m, denom
    if chi is not None and k > chi:
        k = chi
        numer = 0.0
        for v in s[:k]:
            numer += float(v) ** 2
    return k, numer / denom


# ----------------------------------------------------------------------------
# MPS container
# ----------------------------------------------------------------------------

class RefMPS:
    """List of site tensors ``A[s]`` of shape [chi_left, 2, chi_right] (complex128)."""

    def __init__(self, n: int):
        self.n = n
        self.tensors = []
        for _ in range(n):
            t = np.zeros((1, 2, 1), dtype=np.complex128)
            t[0, 0, 0] = 1.0  # |0>, KernelPkg.jl:68 MPS(site_inds, "0")
            self.tensors.append(t)
        # ITensors ortho limits of a product state: llim = 0, rlim = 2 (centre on site 1).
        self.llim = -1          # sites <= llim are left-orthonormal   (0-based)
        self.rlim = 1           # sites >= rlim are right-orthonormal  (0-based)
        self.fidelity = 1.0
        self.n_svd = 0
        self.n_qr = 0

    # -- gauge moves ---------------------------------------------------------
    def orthogonalize(self, j: int) -> None:
        """ITensors ``orthogonalize!(psi, j)``: QR moves, no truncation."""
        while self.llim < j - 1:
            b = self.llim + 1
            a = self.tensors[b]
            cl, _, cr = a.shape
            q, r = np.linalg.qr(a.reshape(cl * 2, cr))
            k = q.shape[1]
            self.tensors[b] = q.reshape(cl, 2, k)
            nxt = self.tensors[b + 1]
            self.tensors[b + 1] = np.tensordot(r, nxt, axes=([1], [0]))
            self.llim = b
            if self.rlim < b + 2:
                self.rlim = b + 2
            self.n_qr += 1
        while self.rlim > j + 1:
            b = self.rlim - 1
            a = self.tensors[b]
            cl, _, cr = a.shape
            q, r = np.linalg.qr(a.reshape(cl, 2 * cr).T)   # a^T = q r  ->  a = r^T q^T
            k = q.shape[1]
            self.tensors[b] = q.T.reshape(k, 2, cr)
            prv = self.tensors[b - 1]
            self.tensors[b - 1] = np.tensordot(prv, r.T, axes=([2], [0]))
            self.rlim = b
            if self.llim > b - 2:
                self.llim = b - 2
            self.n_qr += 1

    # -- gates ---------------------------------------------------------------
    def apply_1q(self, mat: np.ndarray, q: int, orthogonalize: bool = True) -> None:
        if orthogonalize:
            self.orthogonalize(q)
        self.tensors[q] = np.einsum("pq,aqb->apb", mat, self.tensors[q])

    def apply_2q(self, mat: np.ndarray, q: int, cutoff: float, mode: str = "itensors",
                 chi: int | None = None) -> None:
        """Adjacent two-site gate on (q, q+1) followed by SVD truncation."""
        self.orthogonalize(q)
        a, b = self.tensors[q], self.tensors[q + 1]
        cl, cr = a.shape[0], b.shape[2]
        t = np.tensordot(a, b, axes=([2], [0]))                     # [cl, l, r, cr]
        g = mat.reshape(2, 2, 2, 2)                                 # [L, R, l, r]
        theta = np.einsum("LRlr,alrc->aLRc", g, t).reshape(cl * 2, 2 * cr)
        try:
            u, s, vh = np.linalg.svd(theta, full_matrices=False)    # LAPACK gesdd, like ITensors' default
        except np.linalg.LinAlgError:                               # ITensors falls back to gesvd
            import scipy.linalg
            u, s, vh = scipy.linalg.svd(theta, full_matrices=False, lapack_driver="gesvd")
        self.n_svd += 1
        if mode == "itensors":
            k, _ = truncate_itensors(s * s, cutoff, maxdim=chi)
            self.tensors[q] = u[:, :k].reshape(cl, 2, k)
            self.tensors[q + 1] = (s[:k, None] * vh[:k, :]).reshape(k, 2, cr)
            self.llim, self.rlim = q, q + 2
        elif mode == "pytket":
            k, kept = truncate_pytket(s, 1.0 - cutoff, chi=chi)
            sk = s[:k] * np.sqrt(1.0 / kept)
            self.fidelity *= kept
            self.tensors[q] = (u[:, :k] * sk[None, :]).reshape(cl, 2, k)
            self.tensors[q + 1] = vh[:k, :].reshape(k, 2, cr)
            self.llim, self.rlim = q - 1, q + 1
        else:
            raise ValueError(mode)

    # -- queries -------------------------------------------------------------
    def bond_dims(self) -> list[int]:
        return [t.shape[2] for t in self.tensors[:-1]]

    def max_chi(self) -> int:  # maxlinkdim, KernelPkg.jl:70
        return max([1] + self.bond_dims())

    def nbytes(self) -> int:
        return sum(t.nbytes for t in self.tensors)

    def to_statevector(self) -> np.ndarray:
        v = self.tensors[0]
        for t in self.tensors[1:]:
            v = np.tensordot(v, t, axes=([v.ndim - 1], [0]))
        return v.reshape(-1)


def simulate_mps(n_qubits: int, bound_gates, cutoff: float, mode: str = "itensors",
                 chi: int | None = None) -> RefMPS:
    """``build_and_sim_circ`` (KernelPkg.jl:45-72): gate tuples -> truncated MPS."""
    psi = RefMPS(n_qubits)
    for name, qubits, params in bound_gates:
        mat = gate_matrix(name, params[0] if params else None)
        if len(qubits) == 1:
            psi.apply_1q(mat, qubits[0], orthogonalize=(mode == "itensors"))
        else:
            q0, q1 = qubits
            if q1 != q0 + 1:
                raise RuntimeError("two-qubit gates must act on adjacent sites (q, q+1) after routing")
            psi.apply_2q(mat, q0, cutoff, mode=mode, chi=chi)
    return psi


def mps_inner(y: RefMPS, x: RefMPS) -> complex:
    """``inner(y, x) = <y|x>`` (KernelPkg.jl:106)."""
    e = np.ones((1, 1), dtype=np.complex128)
    for ay, ax in zip(y.tensors, x.tensors):
        cy, cx = ay.shape[2], ax.shape[2]
        t = (e @ ax.reshape(ax.shape[0], 2 * cx)).reshape(ay.shape[0] * 2, cx)     # [(a,p), c']
        e = ay.reshape(ay.shape[0] * 2, cy).conj().T @ t                          # [b', c']
    return complex(e[0, 0])


def prepare_for_inner(tensors):
    """Pre-reshaped operands for repeated overlaps (same arithmetic as ``mps_inner``; only the
    per-call reshape / conjugate overhead is hoisted so the CPU baseline is not weaker than ITensors,
    SURVEY.md 8(d))."""
    ket = [np.ascontiguousarray(a.reshape(a.shape[0], 2 * a.shape[2])) for a in tensors]
    bra = [np.ascontiguousarray(a.reshape(a.shape[0] * 2, a.shape[2]).conj().T) for a in tensors]
    dims = [a.shape[2] for a in tensors]
    return ket, bra, dims


def inner_prepared(bra_y, ket_x, dims_x) -> complex:
    e = np.ones((1, 1), dtype=np.complex128)
    for ad, ax, cx in zip(bra_y, ket_x, dims_x):
        e = ad @ (e @ ax).reshape(-1, cx)
    return complex(e[0, 0])
