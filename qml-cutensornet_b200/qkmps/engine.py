"""Host orchestration of the quantum-kernel path on B200s: shard -> simulate -> exchange -> Gram tiles.

One process per GPU.  ``comm`` is duck-typed like the reference's ``mpi_comm`` (Get_rank / Get_size);
with more than one rank it is turned into a ``qkmps.comm.TorchComm`` (torch.distributed: NCCL on GPUs,
gloo for the CPU-side tests of this logic).  PyTorch is used for device buffers, streams and the
collectives only; all arithmetic is in libqkmps.so.

Replaces gpu_backend/kernel_state_ansatz.py:152-428 of the reference:
  * chunking of X over ranks (gpu:154,169-174)            -> ``shard_bounds``
  * per-chunk simulate loop (gpu:213-231,255-273)          -> one stage-1 kernel launch per shard
  * pickled-MPS round robin (gpu:342-352,416-419)          -> one all-gather of the packed states on a side stream,
                                                              overlapped with the local x local block of stage 2
  * ring schedule of chunk pairs (gpu:330-334,366-405)     -> ``panel_tiles``: rank r owns the rows of its shard and
                                                              the column shards r, r-1, .., r-n/2 (cyclic)
  * dense reduce(SUM) to root (gpu:428)                    -> gather of the owned row panels
"""

from __future__ import annotations

import os
import time

import numpy as np

from . import (CHI_LIMIT, DMMA_D_LIMIT, QK_FLAG_CAP_HIT, QK_FLAG_NO_CONVERGE, QkError, frag_stride,
               gram_big, gram_frags, gram_lane, pad_dims, simulate_async, simulate_dev)

# Datapoints per GPU up to which stage 1 uses one CTA cluster per datapoint (B form), and the cluster size.  Measured r2
# (C3 shape, paired distance-2 routing, profiles/r02_stage1_schedules.txt), stage-1 ms sequential | B form by cluster size:
#   125 points:  9.8 | c4 6.6  c5 5.6  c6 5.0  c8 4.8        250 points:  9.8 | c4 8.6  c5 10.5  c6 9.6  c8 8.9
#   500 points: 11.1 | c2 15.5  c4 15.6                      1000 points: 15.0 | c2 27.1  c4 29.1
PARALLEL_MAX_LOCAL = 300   # cluster size: 8 CTAs up to 150 datapoints, 4 above (chosen in qk_api.cu from the batch size)
LANE_CHI_LIMIT = 4   # at or below this bond dimension stage 2 runs one lane per pair on the FP64 CUDA cores
FRAG_D_LIMIT = 16    # padded bond dimension up to which stage 2 runs on the packed-fragment tensor-core kernel; above it the
#                      batched-GEMM sweep on the stores is faster than the CUDA-core fragment kernel (C4 shape at gamma 0.1,
#                      chi <= 17, 32 896 pairs x 165 sites: 48 ms against 119 ms)
# bond caps tried in turn; each has its own kernel configuration: <= 32 shared-memory-resident kernels (one CTA or, in
# B form, one small cluster per datapoint), above that the large-matrix kernel (theta in L2, block Jacobi, one CTA
# cluster per datapoint -- BASELINE config 4)
CAP_LADDER = (4, 8, 16, 24, 32, 64, 128, 256, 512)
BIG_STORE_BUDGET = 24 << 30   # bytes of working store a large-matrix launch may take (its time does not depend on the cap)


class _DevView:
    """Zero-copy uint8 view of raw device memory for torch (CUDA array interface)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2}


def shard_bounds(n_items: int, n_ranks: int, rank: int) -> tuple[int, int]:
    """Contiguous chunk of ceil(N / n_ranks) items per rank (reference gpu:154,169-174)."""
    per = -(-n_items // n_ranks) if n_items > 0 else 0
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)


def panel_tiles(n_x: int, n_y: int, symmetric: bool, n_ranks: int, rank: int):
    """Gram work of ``rank`` as [r0, r1, c0, c1] tiles (rows = y, columns = x).

    Every rank owns the rows of its own shard (the Y shard; the X shard for the symmetric train Gram), so its
    results form one row panel and the matrix is assembled by a gather of panels -- no dense reduce.

    * ``local``: the block of its own rows against its own column shard -- needs no state of another rank, so it
      runs while the exchange is in flight.  Symmetric: only x <= y inside it, mirrored by the kernel.
    * ``remote``: the other column shards.  Rectangular Gram: all of them.  Symmetric Gram: the reference's ring
      schedule (gpu:330-334): column shards r-1, .., r-floor(n/2) (cyclic), all pairs of each block; for an even
      number of ranks the block at distance n/2 is shared with the partner rank (lower rank: first half of its
      own rows x the partner's columns; higher rank: its rows x second half of the partner's rows as columns).
      The matrix entry (y, x) of such a block is stored at [y, x] only; the caller mirrors with max(K, K^T).
    Every unordered pair is computed exactly once.
    """
    if symmetric:
        n_y = n_x
    lo, hi = shard_bounds(n_y, n_ranks, rank)
    out = {"rows": (lo, hi), "local": [], "remote": []}
    if hi <= lo:
        return out
    if not symmetric:
        xlo, xhi = shard_bounds(n_x, n_ranks, rank)
        if xhi > xlo:
            out["local"].append([lo, hi, xlo, xhi])
        if xlo > 0:
            out["remote"].append([lo, hi, 0, xlo])
        if xhi < n_x:
            out["remote"].append([lo, hi, xhi, n_x])
        return out
    out["local"].append([lo, hi, lo, hi])
    for k in range(1, n_ranks // 2 + 1):
        j = (rank - k) % n_ranks
        clo, chi_ = shard_bounds(n_x, n_ranks, j)
        if chi_ <= clo:
            continue
        if n_ranks % 2 == 0 and k == n_ranks // 2:
            if rank < j:      # lower rank of the pair: first half of its own rows
                mid = lo + (hi - lo + 1) // 2
                if mid > lo:
                    out["remote"].append([lo, mid, clo, chi_])
            else:             # higher rank: all its rows x the partner's second half
                mid = clo + (chi_ - clo + 1) // 2
                if chi_ > mid:
                    out["remote"].append([lo, hi, mid, chi_])
        else:
            out["remote"].append([lo, hi, clo, chi_])
    return out


class SingleComm:
    """Stand-in for ``MPI.COMM_WORLD`` when one process drives one GPU (mpi4py is optional)."""

    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise QkError(-2, "no CUDA device: the quantum-kernel path has no CPU fallback")
    return torch


def bond_caps(n_qubits: int, chi_cap: int) -> np.ndarray:
    """Per-bond caps of a plan: min(chi_cap, chain-edge bound 2^min(b, n-b)) (qk_plan.cpp)."""
    b = np.arange(n_qubits + 1)
    e = np.minimum(b, n_qubits - b)
    return np.minimum(chi_cap, 2 ** np.minimum(e, 30)).astype(np.int32)


def store_bytes(n_qubits: int, chi_cap: int) -> int:
    c = bond_caps(n_qubits, chi_cap).astype(np.int64)
    return int((c[:-1] * 2 * c[1:]).sum() * 16)


class ShardStates:
    """The simulated states of one rank's shard: one or more device batches (one per bond-cap level the
    shard needed) plus, for every batch, which of its states are valid and where they sit in the shard."""

    def __init__(self, n_local, n_qubits):
        self.n_local, self.n_qubits = n_local, n_qubits
        self.parts = []            # (Batch, info dict, valid local-in-batch indices, shard positions)
        self.sim_ms = 0.0
        self.launches = 0
        self.cap = 1
        self.schedule = ""
        self.plan = None

    def add(self, batch, info, ok_in_batch, shard_pos):
        self.parts.append((batch, info, np.asarray(ok_in_batch, dtype=np.int64), np.asarray(shard_pos, dtype=np.int64)))

    def info(self):
        n, N = self.n_qubits, self.n_local
        out = dict(chi=np.ones((N, n + 1), dtype=np.int32), fidelity=np.ones(N), trunc_weight=np.zeros(N),
                   nbytes=np.zeros(N, dtype=np.int64), flags=np.zeros(N, dtype=np.int32), sweeps=np.zeros(N, dtype=np.int32),
                   seconds=np.zeros(N))
        for batch, info, ok, pos in self.parts:
            if "seconds" not in info:
                info["seconds"] = batch.unit_seconds()
            for k in out:
                out[k][pos] = info[k][ok]
        return out

    def max_chi(self):
        m = np.ones(self.n_qubits + 1, dtype=np.int32)
        for _, info, ok, _ in self.parts:
            if len(ok):
                m = np.maximum(m, info["chi"][ok].max(axis=0))
        return m

    def pack(self, D, frag_ptr, first_index, stream):
        """All valid states into the frag buffer at positions first_index + shard position."""
        for batch, _, ok, pos in self.parts:
            if not len(ok):
                continue
            dst = np.full(batch.N, -1, dtype=np.int32)
            dst[ok] = first_index + pos
            batch.pack_scatter(D, frag_ptr, dst, stream)
            self.launches += 1

    def single_batch(self):
        """The batch, if one batch holds every state of the shard in shard order."""
        if len(self.parts) == 1:
            b, _, ok, pos = self.parts[0]
            if b.N == self.n_local and np.array_equal(ok, pos) and np.array_equal(ok, np.arange(self.n_local)):
                return b
        return None


def _use_parallel(comm, n_local):
    env = os.environ.get("QK_SCHEDULE", "")
    return (env == "parallel") or (env == "" and comm.Get_size() > 1 and 0 < n_local <= PARALLEL_MAX_LOCAL)


def _to_device(X_shard, device, n_qubits, torch):
    if isinstance(X_shard, torch.Tensor):
        return X_shard.contiguous()
    if len(X_shard):
        return torch.from_numpy(np.ascontiguousarray(X_shard, dtype=np.float64)).to(f"cuda:{device}", non_blocking=False)
    return torch.empty((0, n_qubits), device=f"cuda:{device}", dtype=torch.float64)


def _simulate_shard(plan_factory, X_shard, device, chi_cap, comm, n_qubits, escalate=True):
    """Stage 1 on this rank's shard with per-datapoint bond-cap escalation.

    ``chi_cap`` is the first cap tried (structural bound, clipped, or the user's ``chi``).  Bond dimensions
    are data dependent (SURVEY.md hard part 1): the shard is simulated with ``QK_PLAN_EARLY_EXIT`` (a datapoint
    stops at its first cap hit) and only the datapoints that hit the cap are re-run at the next cap.  Up to 32 the
    caps of ``CAP_LADDER`` are tried one by one (each has its own shared-memory kernel configuration); above 32 the
    large-matrix kernel works on the actual matrix sizes, so a generous cap costs memory, not time, and the ladder
    jumps to the largest cap whose working store fits ``BIG_STORE_BUDGET``.
    ``X_shard``: host numpy array (copied through pinned memory) or CUDA tensor.
    """
    torch = _torch()
    xt = _to_device(X_shard, device, n_qubits, torch)
    torch.cuda.current_stream().synchronize()
    n_local = int(xt.shape[0])
    states = ShardStates(n_local, n_qubits)
    ladder = [c for c in CAP_LADDER if c >= chi_cap]
    if not ladder:
        raise QkError(-3, f"bond dimension cap {chi_cap} above the limit of the stage-1 kernels (chi <= {CHI_LIMIT})")
    stream = torch.cuda.current_stream().cuda_stream

    # Small shards (the multi-GPU regime: 125 datapoints per GPU at 8 GPUs) are bound by the latency of ONE
    # datapoint's op chain.  They run in B form with a cluster of CTAs per datapoint (QK_PLAN_PARALLEL): the
    # dependency depth of the circuit replaces its op count.  Large shards keep the sequential fused schedule,
    # which is throughput-bound anyway (1000 CTAs per 125 datapoints already saturate the SMs).
    parallel = _use_parallel(comm, n_local)
    states.schedule = "parallel (B form, one CTA cluster per datapoint)" if parallel else "sequential (one CTA per datapoint)"

    def run(cap, idx, early):
        plan = plan_factory(cap, early, True) if (parallel and cap <= 32) else plan_factory(cap, early)
        sub = xt if len(idx) == n_local and np.array_equal(idx, np.arange(n_local)) else xt[torch.from_numpy(idx).to(xt.device)]
        batch = simulate_dev(plan, sub.data_ptr(), int(sub.shape[0]), int(sub.shape[1]), device=device, stream=stream)
        info = batch.info()
        hit = (info["flags"] & QK_FLAG_CAP_HIT) != 0
        ok = np.nonzero(~hit)[0]
        states.add(batch, info, ok, idx[ok])
        states.sim_ms += batch.sim_ms()
        states.launches += 1
        states.plan = plan
        return idx[hit]

    pending = np.arange(n_local, dtype=np.int64)
    level = 0
    while len(pending):
        cap = ladder[level]
        if cap > 32:
            # large-matrix kernel: jump to the largest cap whose store fits the budget
            while level + 1 < len(ladder) and len(pending) * store_bytes(n_qubits, ladder[level + 1]) <= BIG_STORE_BUDGET:
                level += 1
            cap = ladder[level]
        last = (level == len(ladder) - 1) or not escalate
        states.cap = max(states.cap, cap)
        if cap > 32:
            states.schedule = "large-matrix kernel (one CTA cluster per datapoint, theta in L2, block Jacobi)"
        pending = run(cap, pending, not last)
        if last and len(pending):
            raise QkError(-3, f"bond dimension exceeds the limit of the stage-1 kernels (chi <= {CHI_LIMIT})")
        level += 1
    if n_local == 0:
        states.plan = plan_factory(ladder[0], False)
    return states


def _as_torch_comm(comm):
    """Multi-rank runs go through torch.distributed; an mpi4py-like communicator is used to bootstrap it."""
    from .comm import TorchComm, init_from_comm
    if comm.Get_size() == 1 or isinstance(comm, TorchComm):
        return comm
    if hasattr(comm, "as_torch_comm"):       # compat/mpi4py stand-in under torchrun
        return comm.as_torch_comm()
    return init_from_comm(comm)                # a real mpi4py communicator: rendezvous through its bcast


def _gather_panels(comm, panel, n_rows, n_cols, per_rows, symmetric, torch):
    """Assemble K on rank 0 from the row panels [per_rows, n_cols] of every rank (replaces reduce(SUM), gpu:428)."""
    from .comm import gather_to_root
    rank, size = comm.Get_rank(), comm.Get_size()
    if size == 1:
        K = panel[:n_rows]
    else:
        full = gather_to_root(comm, panel)
        if rank != 0:
            return None
        K = full.view(size * per_rows, n_cols)[:n_rows]
    if symmetric:
        K = torch.maximum(K, K.t())     # blocks at cyclic distance >= 1 were stored at [y, x] only; entries are >= 0
    return K


class Profile(dict):
    """Profile of one build_gram call.  Entries that need a device read-back (per-state bond dimensions, fidelities,
    per-circuit and per-product times) are produced on first access, so a caller that only wants K pays nothing."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._lazy = {}

    def lazy(self, key, fn):
        self._lazy[key] = fn

    def __missing__(self, key):
        if key in self._lazy:
            self[key] = self._lazy.pop(key)()
            return self[key]
        raise KeyError(key)

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self._lazy


_CLOCK_HZ = {}


def _clock_hz(torch, device):
    if device not in _CLOCK_HZ:
        _CLOCK_HZ[device] = 1e3 * getattr(torch.cuda.get_device_properties(device), "clock_rate", 1.965e6)
    return _CLOCK_HZ[device]


class Checkpoint:
    """Per-rank checkpoint of the Gram row panel (replaces cpu_backend/kernel_state_ansatz.py:212-233,252-253,279-282,326
    of the reference: a per-rank ``tmp/checkpoint_rank_<rank>_<info_file>.npy`` rewritten after every tile, a tile
    counting as done iff its first entry is non-zero).  Here the rank's rows are cut into ``groups`` row groups; after
    each group the panel and an explicit "done" bitmap are written atomically; a restart loads them, skips the finished
    groups and the file is removed once the matrix is complete."""

    def __init__(self, path, groups=4):
        self.path, self.groups = str(path), max(1, int(groups))
        self.groups_run = 0          # row groups computed by this call (tests)
        self.abort_after = None      # tests: raise after this many groups were written

    def _sig(self, shape, rows, n_qubits, symmetric, size):
        return np.array([shape[0], shape[1], rows[0], rows[1], n_qubits, int(symmetric), size, self.groups], dtype=np.int64)

    def load(self, shape, rows, n_qubits, symmetric, size):
        done = np.zeros(self.groups, dtype=bool)
        if os.path.exists(self.path):
            try:
                z = np.load(self.path)
                if np.array_equal(z["sig"], self._sig(shape, rows, n_qubits, symmetric, size)) and z["panel"].shape == tuple(shape):
                    return z["panel"], z["done"].astype(bool)
            except Exception:   # noqa: BLE001  (unreadable / foreign file: start over)
                pass
        return None, done

    def save(self, panel_host, done, rows, n_qubits, symmetric, size):
        os.makedirs(os.path.dirname(self.path) or ".", exist_ok=True)
        tmp = self.path + ".part.npz"
        np.savez(tmp, panel=panel_host, done=done, sig=self._sig(panel_host.shape, rows, n_qubits, symmetric, size))
        os.replace(tmp, self.path)

    def remove(self):
        if os.path.exists(self.path):
            os.remove(self.path)


def build_gram(comm, plan_factory, n_qubits, X, Y=None, chi_cap=16, device=None, return_device=False,
               structural_cap=False, checkpoint=None):
    """Full path.  Returns (K on rank 0 / None elsewhere, profile dict).

    ``X`` / ``Y``: host numpy arrays, or CUDA float64 tensors already resident in HBM (every rank
    holds the full arrays, as in the reference).  ``return_device`` leaves K on the GPU (rank 0).
    ``structural_cap``: ``chi_cap`` is at least the structural bound of the ansatz, so no state can ask for more:
    the whole step is then queued without a host round trip (streamlined mode).
    """
    torch = _torch()
    comm = _as_torch_comm(comm)
    rank, size = comm.Get_rank(), comm.Get_size()
    if device is None:
        device = rank % torch.cuda.device_count()   # reference gpu:152
    torch.cuda.set_device(device)
    if not isinstance(X, torch.Tensor):
        X = np.asarray(X, dtype=np.float64)
    if Y is not None and not isinstance(Y, torch.Tensor):
        Y = np.asarray(Y, dtype=np.float64)
    streamlined = (structural_cap and 4 < chi_cap <= DMMA_D_LIMIT and os.environ.get("QK_ENGINE", "") != "general"
                   and os.environ.get("QK_GRAM_BIG", "0") != "1" and checkpoint is None)
    if streamlined:
        return _build_gram_streamlined(comm, plan_factory, n_qubits, X, Y, chi_cap, device, return_device, torch)
    return _build_gram_general(comm, plan_factory, n_qubits, X, Y, chi_cap, device, return_device, torch, checkpoint)


def _to_host(K, torch):
    """Device -> host through pinned memory (torch's pinned allocator caches the buffer between calls): 8 MB in
    ~0.35 ms instead of ~1.5 ms through pageable memory."""
    host = torch.empty(K.shape, dtype=K.dtype, pin_memory=True)
    host.copy_(K)
    return host.numpy()


def _tile_clock_stats(clk, used, torch, device):
    """Per-inner-product seconds from the per-CTA-tile clocks of the tensor-core kernel (8 pairs per tile)."""
    if clk is None or used <= 0:
        return None
    ticks = clk[:used].cpu().numpy().astype(np.float64)
    ticks = ticks[ticks > 0]
    if not len(ticks):
        return None
    return ticks / _clock_hz(torch, device) / 8.0


def _build_gram_streamlined(comm, plan_factory, n_qubits, X, Y, chi_cap, device, return_device, torch):
    """No state can exceed ``chi_cap``: stage 1 -> pack -> [all-gather on a side stream || local block] -> remote
    blocks -> gather of panels, all queued on the device without a host synchronisation in between."""
    from .comm import allgather_into
    rank, size = comm.Get_rank(), comm.Get_size()
    dev = f"cuda:{device}"
    symmetric = Y is None
    Nx = len(X)
    Ny = Nx if symmetric else len(Y)
    nb = n_qubits + 1
    t_all = time.perf_counter()
    trace = []

    def mark(name):
        trace.append((name, (time.perf_counter() - t_all) * 1e3))
    main = torch.cuda.current_stream()
    stream = main.cuda_stream
    prof = {}

    lo, hi = shard_bounds(Nx, size, rank)
    parallel = _use_parallel(comm, hi - lo)
    plan = plan_factory(chi_cap, False, True) if parallel else plan_factory(chi_cap, False)
    D = pad_dims(bond_caps(n_qubits, chi_cap))
    stride = frag_stride(n_qubits, D)
    per_x = -(-Nx // size)
    fragX = torch.empty(size * max(per_x, 1) * stride, dtype=torch.uint8, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    mark("alloc")
    xt = _to_device(X[lo:hi], device, n_qubits, torch)
    bx = simulate_async(plan, xt.data_ptr(), int(xt.shape[0]), int(xt.shape[1]) if xt.shape[0] else n_qubits,
                        device=device, stream=stream)
    mark("simulate queued")
    bx.pack_async(D, fragX.data_ptr(), rank * per_x, stream)
    bx.release_store(stream)        # only the packed fragments are needed from here on
    mark("pack queued")
    launches = 2
    by, fragY, ylo, yhi = None, None, lo, hi
    if not symmetric:
        ylo, yhi = shard_bounds(Ny, size, rank)
        yt = _to_device(Y[ylo:yhi], device, n_qubits, torch)
        by = simulate_async(plan, yt.data_ptr(), int(yt.shape[0]), int(yt.shape[1]) if yt.shape[0] else n_qubits,
                            device=device, stream=stream)
        fragY = torch.empty(max(yhi - ylo, 1) * stride, dtype=torch.uint8, device=dev)   # bras are only needed locally
        by.pack_async(D, fragY.data_ptr(), 0, stream)
        by.release_store(stream)
        launches += 2
    gathered = None
    if size > 1:
        packed = torch.cuda.Event()
        packed.record(main)
        side = torch.cuda.Stream(device=device)
        side.wait_event(packed)
        with torch.cuda.stream(side):
            allgather_into(comm, fragX, rank * per_x * stride, per_x * stride)
            gathered = torch.cuda.Event()
            gathered.record(side)

    work = panel_tiles(Nx, Ny, symmetric, size, rank)
    per_rows = -(-Ny // size)
    panel = torch.zeros((max(per_rows, 1), Nx), dtype=torch.float64, device=dev)
    k_ptr = panel.data_ptr() - work["rows"][0] * Nx * 8          # the kernels index rows globally
    fy_ptr = None if symmetric else fragY.data_ptr() - ylo * stride
    n_tiles_max = ((per_rows + 1) // 2 + 2) * ((Nx + 3) // 4 + 2)
    clk = torch.zeros(n_tiles_max, dtype=torch.int64, device=dev) if size == 1 and os.environ.get("QK_TILE_CLOCKS", "1") != "0" else None
    ev[0].record(main)
    if work["local"]:
        # the local block reads only this rank's states: it runs while the all-gather is in flight
        gram_frags(device, n_qubits, D, fragX.data_ptr(), hi, None if symmetric else D, fy_ptr, yhi,
                   work["local"], symmetric, k_ptr, Nx, stream, wait=False,
                   tile_clocks=(clk.data_ptr(), n_tiles_max) if clk is not None else None)
        launches += 1
    mark("local gram queued")
    ev[1].record(main)
    if gathered is not None:
        main.wait_event(gathered)
    ev[2].record(main)
    if work["remote"]:
        gram_frags(device, n_qubits, D, fragX.data_ptr(), Nx, D, fragX.data_ptr() if symmetric else fy_ptr,
                   Nx if symmetric else yhi, work["remote"], False, k_ptr, Nx, stream, wait=False)
        launches += 1
    ev[3].record(main)
    mark("remote gram queued")
    K = _gather_panels(comm, panel, Ny, Nx, max(per_rows, 1), symmetric, torch)
    mark("panels gathered (queued)")
    out = None
    if rank == 0:
        out = K if return_device else _to_host(K, torch)
    torch.cuda.synchronize()
    mark("device idle")

    # correctness first: one small read-back of the stage-1 flags (anything else of the profile is read on demand)
    for bb in (bx, by):
        if bb is not None and bb.N:
            fl = bb.flags_or()
            if fl & QK_FLAG_NO_CONVERGE:
                raise QkError(-2, "stage 1: the Jacobi SVD hit its sweep limit on at least one state")
            if fl & QK_FLAG_CAP_HIT:
                raise QkError(-3, "a state exceeded a bond cap that was declared structural")
    mark("flags checked")

    def full_info(bb):
        inf = bb.info()
        inf["seconds"] = bb.unit_seconds()
        return inf
    prof = Profile(prof)
    prof.lazy("info_x", lambda: full_info(bx))
    prof.lazy("info_y", (lambda: full_info(by)) if by is not None else (lambda: None))
    prof["_keep"] = (bx, by, clk)
    prof["sim_ms_x"] = bx.sim_ms()
    prof["sim_ms_y"] = by.sim_ms() if by is not None else 0.0
    prof["chi_cap"] = chi_cap
    prof["plan"] = plan.info()
    prof["plan_obj"] = plan
    prof["stage1_schedule"] = ("parallel (B form, one CTA cluster per datapoint)" if parallel
                               else "sequential (one CTA per datapoint)")
    prof["shard"] = (lo, hi)
    prof["gram_kernel"] = "qk_gram_dmma_kernel"
    prof["gram_ms_local"] = ev[0].elapsed_time(ev[1])
    prof["gram_ms_remote"] = ev[2].elapsed_time(ev[3])
    prof["gram_ms"] = prof["gram_ms_local"] + prof["gram_ms_remote"]
    prof["exchange_wait_ms"] = ev[1].elapsed_time(ev[2])      # what the all-gather costs beyond the local block
    prof["exchange_s"] = prof["exchange_wait_ms"] * 1e-3
    prof["exchange"] = ("none (one rank)" if size == 1 else
                        "all-gather of packed states on a side stream, overlapped with the local x local block; "
                        "row panels gathered to rank 0")
    prof["frag_bytes_per_state"] = (stride, stride)
    prof["Dx"], prof["Dy"] = D, D
    prof["launches"] = launches
    prof["mode"] = "streamlined"
    mark("events")
    prof.lazy("pair_seconds", lambda: _tile_clock_stats(clk, n_tiles_max, torch, device))
    prof["no_converge"] = 0
    mark("profile read back")
    prof["host_trace_ms"] = trace
    prof["total_s"] = time.perf_counter() - t_all
    return out, prof


def _build_gram_general(comm, plan_factory, n_qubits, X, Y, chi_cap, device, return_device, torch, checkpoint=None):
    """Data-dependent bond dimensions: stage 1 with cap escalation, then the stage-2 path is chosen from the measured
    (all-reduced) bond dimensions -- lane-per-pair kernel (chi <= 4), tensor-core fragments (padded D <= 16),
    CUDA-core fragments (D <= 32) or batched GEMMs on the stores (any chi)."""
    from .comm import allgather_bytes, allreduce_max_array
    rank, size = comm.Get_rank(), comm.Get_size()
    dev = f"cuda:{device}"
    symmetric = Y is None
    Nx = len(X)
    Ny = Nx if symmetric else len(Y)
    prof = {}
    t_all = time.perf_counter()
    launches = 0

    # ---- stage 1 on this rank's shard(s).  A failure here (bond dimension above the kernels' limit, CUDA error) is
    # rank-local and data dependent: it is carried through the collective below and raised on EVERY rank, so that
    # the other ranks do not block forever in all_reduce / all_gather.
    lo, hi = shard_bounds(Nx, size, rank)
    ylo, yhi = (lo, hi) if symmetric else shard_bounds(Ny, size, rank)
    sx, sy, info_x, info_y, local_err = None, None, None, None, None
    try:
        sx = _simulate_shard(plan_factory, X[lo:hi], device, chi_cap, comm, n_qubits)
        info_x = sx.info()
        if not symmetric:
            sy = _simulate_shard(plan_factory, Y[ylo:yhi], device, chi_cap, comm, n_qubits)
            info_y = sy.info()
        for inf in (info_x, info_y):
            if inf is not None and len(inf["flags"]) and np.any(inf["flags"] & QK_FLAG_NO_CONVERGE):
                raise QkError(-2, "stage 1: the Jacobi SVD hit its sweep limit on at least one state")
    except QkError as e:
        if size == 1:
            raise
        local_err = e
    if size > 1:
        codes = allreduce_max_array(comm, np.array([0 if local_err is None else -int(local_err.code)], dtype=np.int32))
        if int(codes[0]) != 0:
            if local_err is not None:
                raise local_err
            raise QkError(-int(codes[0]), "stage 1 failed on another rank (see that rank's error)")
    prof["sim_ms_x"] = sx.sim_ms
    prof["sim_ms_y"] = sy.sim_ms if sy is not None else 0.0
    prof["chi_cap"] = max(sx.cap, sy.cap if sy is not None else 1)
    prof["info_x"], prof["info_y"] = info_x, info_y
    prof["plan"] = sx.plan.info()
    prof["stage1_schedule"] = sx.schedule
    prof["plan_obj"] = sx.plan
    prof["shard"] = (lo, hi)
    prof["mode"] = "general"

    # ---- exchange: one collective for the per-bond maxima, "plain batch" flag and the common bond cap
    t0 = time.perf_counter()
    lane_local = sx.single_batch() is not None or sx.n_local == 0
    if not symmetric:
        lane_local = lane_local and (sy.single_batch() is not None or sy.n_local == 0) and \
            (sx.n_local == 0 or sy.n_local == 0 or sy.cap == sx.cap)
    flag = np.array([0 if lane_local else 1, max(sx.cap, sy.cap if sy is not None else 1)], dtype=np.int32)
    red = allreduce_max_array(comm, np.concatenate([sx.max_chi(), sy.max_chi() if not symmetric else sx.max_chi(), flag]))
    nb = n_qubits + 1
    Dx = pad_dims(red[:nb])
    Dy = Dx if symmetric else pad_dims(red[nb:2 * nb])
    max_chi = int(red[:2 * nb].max())
    # layout of the exchanged stores: the smallest cap of the ladder that covers the measured bond dimensions (a shard
    # may have been simulated with a far larger cap -- above 32 a generous cap costs stage 1 nothing, but the
    # exchanged buffer is sized by it)
    cap_common = next((c for c in CAP_LADDER if c >= max_chi), int(red[-1]))
    use_lane = (max_chi <= LANE_CHI_LIMIT and int(red[-2]) == 0 and os.environ.get("QK_GRAM_LANE", "1") != "0")
    use_big = (not use_lane) and (int(max(Dx.max(), Dy.max())) > FRAG_D_LIMIT or os.environ.get("QK_GRAM_BIG", "0") == "1")
    stream = torch.cuda.current_stream().cuda_stream
    work = panel_tiles(Nx, Ny, symmetric, size, rank)
    per_x, per_rows = -(-Nx // size), -(-Ny // size)
    panel = torch.zeros((max(per_rows, 1), Nx), dtype=torch.float64, device=dev)
    k_ptr = panel.data_ptr() - work["rows"][0] * Nx * 8          # the kernels index rows globally
    ms = 0.0
    prof["exchange"] = "none (one rank)" if size == 1 else "all-gather of the ket states; bras stay local; row panels gathered to rank 0"

    def launches_for(run):
        """local block (x <= y mirrored when symmetric), then the remote blocks (all pairs, stored at [y, x] only);
        with a checkpoint: one row group of this rank's panel at a time, written to disk after each"""
        if checkpoint is None:
            t = 0.0
            if work["local"]:
                t += run(work["local"], symmetric, True)
            if work["remote"]:
                t += run(work["remote"], False, False)
            return t
        r_lo, r_hi = work["rows"]
        saved, done = checkpoint.load(tuple(panel.shape), (r_lo, r_hi), n_qubits, symmetric, size)
        if saved is not None:
            panel.copy_(torch.from_numpy(saved).to(dev))
        step = -(-max(r_hi - r_lo, 1) // checkpoint.groups)
        t = 0.0
        for g in range(checkpoint.groups):
            g0, g1 = r_lo + g * step, min(r_lo + (g + 1) * step, r_hi)
            if done[g] or g1 <= g0:
                done[g] = True
                continue

            def clip(tiles):
                out = []
                for r0, r1, c0, c1 in tiles:
                    a, b = max(r0, g0), min(r1, g1)
                    if b > a and c1 > c0:
                        out.append([a, b, c0, c1])
                return out
            loc, rem = clip(work["local"]), clip(work["remote"])
            if loc:
                t += run(loc, symmetric, True)
            if rem:
                t += run(rem, False, False)
            torch.cuda.synchronize()
            done[g] = True
            checkpoint.groups_run += 1
            checkpoint.save(panel.cpu().numpy(), done, (r_lo, r_hi), n_qubits, symmetric, size)
            if checkpoint.abort_after is not None and checkpoint.groups_run >= checkpoint.abort_after:
                raise KeyboardInterrupt("checkpoint test: interrupted after %d row groups" % checkpoint.groups_run)
        return t

    if use_lane or use_big:
        # the unpadded stores themselves (+ bond dimensions) are exchanged
        if use_lane:
            prof["gram_kernel"] = "qk_gram_lane_kernel"
            plan = sx.plan
        else:
            # every rank re-lays its states out for the common bond cap (shards may have escalated differently)
            prof["gram_kernel"] = "qk_big_gemm_kernel"
            plan = plan_factory(cap_common, False)
        stride_b = int(plan.info().state_stride) * 16

        def stores(shard, n_total, gather):
            """(store ptr, chi ptr, keep-alive) addressed by GLOBAL state index"""
            first = shard_bounds(n_total, size, rank)[0]
            b = shard.single_batch()
            if use_lane and size == 1:
                ptr, _, chi_ptr = b.store()
                return ptr, chi_ptr, None
            per = -(-n_total // size)
            n_slots = size * max(per, 1) if gather else max(shard.n_local, 1)
            base = rank * per if gather else 0
            st = torch.empty(n_slots * stride_b, dtype=torch.uint8, device=dev)
            ch = torch.ones(n_slots * nb, dtype=torch.int32, device=dev)
            if use_lane:
                if b is not None and b.N:
                    ptr, _, chi_ptr = b.store()
                    st[base * stride_b:(base + b.N) * stride_b].copy_(torch.as_tensor(_DevView(ptr, b.N * stride_b), device=dev))
                    ch[base * nb:(base + b.N) * nb].copy_(torch.as_tensor(_DevView(chi_ptr, b.N * nb * 4), device=dev).view(torch.int32))
            else:
                for batch, _, ok, pos in shard.parts:
                    if len(ok):
                        dst = np.full(batch.N, -1, dtype=np.int32)
                        dst[ok] = base + pos
                        batch.repack(plan, st.data_ptr(), ch.data_ptr(), dst, stream)
                        shard.launches += 1
            if gather and size > 1:
                from .comm import allgather_into
                allgather_into(comm, st, rank * per * stride_b, per * stride_b)
                allgather_into(comm, ch.view(torch.uint8), rank * per * nb * 4, per * nb * 4)
                return st.data_ptr(), ch.data_ptr(), (st, ch)
            off = 0 if gather else first
            return st.data_ptr() - off * stride_b, ch.data_ptr() - off * nb * 4, (st, ch)

        px, cx, keep_x = stores(sx, Nx, True)
        py, cy, keep_y = (px, cx, None) if symmetric else stores(sy, Ny, False)
        torch.cuda.synchronize()
        prof["exchange_s"] = time.perf_counter() - t0
        prof["frag_bytes_per_state"] = (stride_b, stride_b)

        def run(tiles, sym, is_local):
            nonlocal launches
            ny_arg = Nx if symmetric else yhi
            if use_lane:
                launches += 1
                return gram_lane(plan, device, max_chi, px, cx, Nx, None if sym else py, None if sym else cy, ny_arg,
                                 tiles, sym, k_ptr, Nx, stream)
            launches += 2 * n_qubits + 2
            return gram_big(plan, device, red[:nb], None if sym else red[(0 if symmetric else nb):(nb if symmetric else 2 * nb)],
                            px, cx, Nx, None if sym else py, None if sym else cy, ny_arg, tiles, sym, k_ptr, Nx, stream)

        ms = launches_for(run)
        del keep_x, keep_y
    else:
        prof["gram_kernel"] = "qk_gram_dmma_kernel" if int(max(Dx.max(), Dy.max())) <= DMMA_D_LIMIT else \
            "qk_gram_frag_generic_kernel"   # D > 16: CUDA-core kernel on the same packed buffers (any rank count)
        stride_x = frag_stride(n_qubits, Dx)
        fx = torch.empty(size * max(per_x, 1) * stride_x, dtype=torch.uint8, device=dev)
        sx.pack(Dx, fx.data_ptr(), rank * per_x, stream)
        if size > 1:
            from .comm import allgather_into
            allgather_into(comm, fx, rank * per_x * stride_x, per_x * stride_x)
        stride_y, fy_ptr = stride_x, fx.data_ptr()
        fy = None
        if not symmetric:
            stride_y = frag_stride(n_qubits, Dy)
            fy = torch.empty(max(sy.n_local, 1) * stride_y, dtype=torch.uint8, device=dev)    # bras are only needed locally
            sy.pack(Dy, fy.data_ptr(), 0, stream)
            fy_ptr = fy.data_ptr() - ylo * stride_y
        torch.cuda.synchronize()
        prof["exchange_s"] = time.perf_counter() - t0
        prof["frag_bytes_per_state"] = (stride_x, stride_y)

        def run(tiles, sym, is_local):
            nonlocal launches
            launches += 1
            return gram_frags(device, n_qubits, Dx, fx.data_ptr(), Nx, None if sym else Dy, None if sym else fy_ptr,
                              Nx if symmetric else yhi, tiles, sym, k_ptr, Nx, stream)

        ms = launches_for(run)
        del fy
    launches += sx.launches + (sy.launches if sy is not None else 0)
    prof["gram_ms"] = ms
    prof["Dx"], prof["Dy"] = Dx, Dy
    prof["launches"] = launches
    K = _gather_panels(comm, panel, Ny, Nx, max(per_rows, 1), symmetric, torch)
    out = None
    if rank == 0:
        out = K if return_device else _to_host(K, torch)
    torch.cuda.synchronize()
    if checkpoint is not None:
        checkpoint.remove()          # complete: like the reference (cpu:326)
    prof["pair_seconds"] = None
    prof["total_s"] = time.perf_counter() - t_all
    prof["no_converge"] = 0   # a sweep-limit hit raises above (on every rank)
    return out, prof
