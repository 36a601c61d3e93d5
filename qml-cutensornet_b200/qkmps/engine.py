"""Host orchestration of the quantum-kernel path on B200s: shard -> simulate -> exchange -> Gram tiles.

One process per GPU.  ``comm`` is duck-typed like the reference's ``mpi_comm`` (Get_rank / Get_size);
with more than one rank it must be a ``qkmps.comm.TorchComm`` (torch.distributed: NCCL on GPUs,
gloo for the CPU-side tests of this logic).  PyTorch is used for device buffers, streams and the
collectives only; all arithmetic is in libqkmps.so.

Replaces gpu_backend/kernel_state_ansatz.py:152-428 of the reference:
  * chunking of X over ranks (gpu:154,169-174)            -> ``shard_bounds``
  * per-chunk simulate loop (gpu:213-231,255-273)          -> one stage-1 kernel launch per shard
  * pickled-MPS round robin (gpu:342-352,416-419)          -> one all-gather of the packed "frag" buffers
  * per-pair vdot loop + symmetric fill (gpu:366-405)      -> stage-2 kernel over this rank's row blocks
  * dense reduce(SUM) to root (gpu:428)                    -> reduce of the (disjointly filled) K buffers
"""

from __future__ import annotations

import time

import numpy as np

import os

from . import (CHI_LIMIT, DMMA_D_LIMIT, QK_FLAG_CAP_HIT, QK_FLAG_NO_CONVERGE, QkError, frag_stride,
               gram_big, gram_frags, gram_lane, pad_dims, simulate_dev)

PARALLEL_MAX_LOCAL = 150   # datapoints per GPU up to which stage 1 uses one CTA cluster per datapoint
LANE_CHI_LIMIT = 4   # at or below this bond dimension stage 2 runs one lane per pair on the FP64 CUDA cores


class _DevView:
    """Zero-copy uint8 view of raw device memory for torch (CUDA array interface)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2}


ROW_BLOCK = 8   # Gram rows are dealt to ranks in blocks of this many rows (multiple of the kernel's TJ)


def shard_bounds(n_items: int, n_ranks: int, rank: int) -> tuple[int, int]:
    """Contiguous chunk of ceil(N / n_ranks) items per rank (reference gpu:154,169-174)."""
    per = -(-n_items // n_ranks) if n_items > 0 else 0
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)


def row_tiles(n_rows: int, n_cols: int, symmetric: bool, n_ranks: int, rank: int, row_block: int = ROW_BLOCK):
    """Row blocks owned by ``rank`` as [r0, r1, c0, c1] tiles.

    Symmetric (train) Gram: only x <= y is computed, so row block b costs ~ (b+1) units; blocks are
    dealt in a boustrophedon (0..R-1, R-1..0, ...) order, which balances the triangle to within one
    block per rank.  Rectangular Gram: plain cyclic deal.
    """
    tiles = []
    n_blocks = -(-n_rows // row_block)
    for b in range(n_blocks):
        rnd, pos = divmod(b, n_ranks)
        owner = pos if (rnd % 2 == 0 or not symmetric) else n_ranks - 1 - pos
        if owner != rank:
            continue
        r0, r1 = b * row_block, min((b + 1) * row_block, n_rows)
        c1 = min(r1, n_cols) if symmetric else n_cols
        if tiles and tiles[-1][1] == r0:
            # contiguous with the previous block of this rank (always, on one rank): one larger tile, so that the
            # L2 supertiles of qk_gram_frags can block over rows as well as columns
            tiles[-1][1], tiles[-1][3] = r1, c1
        else:
            tiles.append([r0, r1, 0, c1])
    return tiles


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


# bond caps tried in turn; each has its own kernel configuration: <= 32 shared-memory-resident kernels (one CTA or, in
# B form, one small cluster per datapoint), above that the large-matrix kernel (theta in L2, block Jacobi, one CTA
# cluster per datapoint -- BASELINE config 4)
CAP_LADDER = (4, 8, 16, 24, 32, 64, 128, 256)
FRAG_D_LIMIT = 32    # padded bond dimension up to which stage 2 runs on the packed-fragment kernels


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

    def add(self, batch, info, ok_in_batch, shard_pos):
        self.parts.append((batch, info, np.asarray(ok_in_batch, dtype=np.int64), np.asarray(shard_pos, dtype=np.int64)))

    def info(self):
        n, N = self.n_qubits, self.n_local
        out = dict(chi=np.ones((N, n + 1), dtype=np.int32), fidelity=np.ones(N), trunc_weight=np.zeros(N),
                   nbytes=np.zeros(N, dtype=np.int64), flags=np.zeros(N, dtype=np.int32), sweeps=np.zeros(N, dtype=np.int32))
        for _, info, ok, pos in self.parts:
            for k in out:
                out[k][pos] = info[k][ok]
        return out

    def max_chi(self):
        m = np.ones(self.n_qubits + 1, dtype=np.int32)
        for _, info, ok, _ in self.parts:
            if len(ok):
                m = np.maximum(m, info["chi"][ok].max(axis=0))
        return m

    def pack(self, D, frag_ptr, stream):
        for batch, _, ok, pos in self.parts:
            if not len(ok):
                continue
            dst = np.full(batch.N, -1, dtype=np.int32)
            dst[ok] = pos
            batch.pack_scatter(D, frag_ptr, dst, stream)
            self.launches += 1

    def single_batch(self):
        """The batch, if one batch holds every state of the shard in shard order (CUDA-core Gram fallback)."""
        if len(self.parts) == 1:
            b, _, ok, pos = self.parts[0]
            if b.N == self.n_local and np.array_equal(ok, pos) and np.array_equal(ok, np.arange(self.n_local)):
                return b
        return None


def _simulate_shard(plan_factory, X_shard, device, chi_cap, comm, n_qubits, escalate=True):
    """Stage 1 on this rank's shard with per-datapoint bond-cap escalation.

    ``chi_cap`` is the first cap tried (structural bound, clipped, or the user's ``chi``).  Bond dimensions
    are data dependent (SURVEY.md hard part 1): the shard is simulated with ``QK_PLAN_EARLY_EXIT`` (a datapoint
    stops at its first cap hit) and only the datapoints that hit the cap are re-run at the next cap of
    ``CAP_LADDER``.  (Starting below the structural bound does not pay: at small bond dimension the kernel is
    bound by the per-op latency of one datapoint, not by its shared-memory footprint -- measured, DESIGN.md.)
    ``X_shard``: host numpy array (copied through pinned memory) or CUDA tensor.
    """
    torch = _torch()
    if isinstance(X_shard, torch.Tensor):
        xt = X_shard.contiguous()
    elif len(X_shard):
        xt = torch.from_numpy(np.ascontiguousarray(X_shard, dtype=np.float64)).to(f"cuda:{device}")
    else:
        xt = torch.empty((0, n_qubits), device=f"cuda:{device}", dtype=torch.float64)
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
    env = os.environ.get("QK_SCHEDULE", "")
    parallel = (env == "parallel") or (env == "" and comm.Get_size() > 1 and 0 < n_local <= PARALLEL_MAX_LOCAL)

    states.schedule = "parallel (B form, one CTA cluster per datapoint)" if parallel else "sequential (one CTA per datapoint)"

    def run(cap, idx, early):
        plan = plan_factory(cap, early, True) if parallel else plan_factory(cap, early)
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
        last = (level == len(ladder) - 1) or not escalate
        states.cap = max(states.cap, cap)
        pending = run(cap, pending, not last)
        if last and len(pending):
            raise QkError(-3, f"bond dimension exceeds the limit of the stage-1 kernels (chi <= {CHI_LIMIT})")
        level += 1
    if n_local == 0:
        states.plan = plan_factory(ladder[0], False)
    return states


def build_gram(comm, plan_factory, n_qubits, X, Y=None, chi_cap=16, device=None, return_device=False):
    """Full path.  Returns (K on rank 0 / None elsewhere, profile dict).

    ``X`` / ``Y``: host numpy arrays, or CUDA float64 tensors already resident in HBM (every rank
    holds the full arrays, as in the reference).  ``return_device`` leaves K on the GPU (rank 0).
    """
    torch = _torch()
    from .comm import allgather_bytes, allreduce_max_array, reduce_sum_to_root
    rank, size = comm.Get_rank(), comm.Get_size()
    if device is None:
        device = rank % torch.cuda.device_count()   # reference gpu:152
    torch.cuda.set_device(device)
    dev = f"cuda:{device}"
    if not isinstance(X, torch.Tensor):
        X = np.asarray(X, dtype=np.float64)
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
    sy, info_y, local_err = None, None, None
    try:
        sx = _simulate_shard(plan_factory, X[lo:hi], device, chi_cap, comm, n_qubits)
        info_x = sx.info()
        if not symmetric:
            if not isinstance(Y, torch.Tensor):
                Y = np.asarray(Y, dtype=np.float64)
            ylo, yhi = shard_bounds(Ny, size, rank)
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

    # ---- exchange: batch-uniform padded dims, pack, all-gather
    t0 = time.perf_counter()
    # one collective for the per-bond maxima and for "every rank holds its shard as one plain batch"
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
    cap_common = int(red[-1])
    use_lane = (max_chi <= LANE_CHI_LIMIT and int(red[-2]) == 0 and os.environ.get("QK_GRAM_LANE", "1") != "0")
    use_big = (not use_lane) and (int(max(Dx.max(), Dy.max())) > FRAG_D_LIMIT or os.environ.get("QK_GRAM_BIG", "0") == "1")
    stream = torch.cuda.current_stream().cuda_stream
    K = torch.zeros((Ny, Nx), dtype=torch.float64, device=dev)
    tiles = row_tiles(Ny, Nx, symmetric, size, rank)
    ms = 0.0

    if use_lane:
        # bond dimensions <= 4: the unpadded stores themselves are exchanged and read by the lane-per-pair kernel
        prof["gram_kernel"] = "qk_gram_lane_kernel"
        plan = sx.plan
        stride_b = int(plan.info().state_stride) * 16

        def gathered(shard, n_total):
            b = shard.single_batch()
            if size == 1:
                ptr, _, chi_ptr = b.store()
                return ptr, chi_ptr, None
            per = -(-n_total // size)
            st = torch.zeros(max(per, 1) * stride_b, dtype=torch.uint8, device=dev)
            ch = torch.ones(max(per, 1) * nb, dtype=torch.int32, device=dev)
            if b is not None and b.N:
                ptr, _, chi_ptr = b.store()
                st[:b.N * stride_b].copy_(torch.as_tensor(_DevView(ptr, b.N * stride_b), device=dev))
                ch[:b.N * nb].copy_(torch.as_tensor(_DevView(chi_ptr, b.N * nb * 4), device=dev).view(torch.int32))
            st_all, ch_all = allgather_bytes(comm, st), allgather_bytes(comm, ch)
            return st_all.data_ptr(), ch_all.data_ptr(), (st_all, ch_all)

        px, cx, keep_x = gathered(sx, Nx)
        py, cy, keep_y = (px, cx, None) if symmetric else gathered(sy, Ny)
        torch.cuda.synchronize()
        prof["exchange_s"] = time.perf_counter() - t0
        prof["frag_bytes_per_state"] = (stride_b, stride_b)
        if tiles:
            ms = gram_lane(plan, device, max_chi, px, cx, Nx, None if symmetric else py, None if symmetric else cy, Ny,
                           tiles, symmetric, K.data_ptr(), Nx, stream)
            launches += 1
        del keep_x, keep_y
    elif use_big:
        # bond dimensions above the fragment kernels (BASELINE config 4): every rank re-lays its states out for the
        # common bond cap, the unpadded stores + bond dimensions are exchanged, and the transfer sweep runs as
        # batched complex GEMMs on the FP64 tensor cores
        prof["gram_kernel"] = "qk_big_gemm_kernel"
        plan = plan_factory(cap_common, False)
        stride_b = int(plan.info().state_stride) * 16

        def relaid(shard, n_total):
            per = -(-n_total // size)
            st = torch.empty(max(per, 1) * stride_b, dtype=torch.uint8, device=dev)
            ch = torch.ones(max(per, 1) * nb, dtype=torch.int32, device=dev)
            for batch, _, ok, pos in shard.parts:
                if len(ok):
                    dst = np.full(batch.N, -1, dtype=np.int32)
                    dst[ok] = pos
                    batch.repack(plan, st.data_ptr(), ch.data_ptr(), dst, stream)
                    shard.launches += 1
            if size == 1:
                return st, ch
            return allgather_bytes(comm, st), allgather_bytes(comm, ch)

        stx, chx = relaid(sx, Nx)
        sty, chy = (stx, chx) if symmetric else relaid(sy, Ny)
        torch.cuda.synchronize()
        prof["exchange_s"] = time.perf_counter() - t0
        prof["frag_bytes_per_state"] = (stride_b, stride_b)
        if tiles:
            ms = gram_big(plan, device, red[:nb], None if symmetric else red[nb:2 * nb], stx.data_ptr(), chx.data_ptr(), Nx,
                          None if symmetric else sty.data_ptr(), None if symmetric else chy.data_ptr(), Ny, tiles,
                          symmetric, K.data_ptr(), Nx, stream)
            launches += 2 * n_qubits + 2
        del stx, chx, sty, chy
    else:
        prof["gram_kernel"] = "qk_gram_dmma_kernel" if int(max(Dx.max(), Dy.max())) <= DMMA_D_LIMIT else \
            "qk_gram_frag_generic_kernel"   # D > 16: CUDA-core kernel on the same packed buffers (any rank count)

        def packed(shard, D, n_total):
            stride = frag_stride(n_qubits, D)
            per = -(-n_total // size)
            local = torch.empty(max(per, 1) * stride, dtype=torch.uint8, device=dev)
            shard.pack(D, local.data_ptr(), stream)
            if size == 1:
                return local, stride
            return allgather_bytes(comm, local), stride

        fx, stride_x = packed(sx, Dx, Nx)
        fy, stride_y = (fx, stride_x) if symmetric else packed(sy, Dy, Ny)
        torch.cuda.synchronize()
        prof["exchange_s"] = time.perf_counter() - t0
        prof["frag_bytes_per_state"] = (stride_x, stride_y)

        # ---- stage 2 on this rank's row blocks
        if tiles:
            ms = gram_frags(device, n_qubits, Dx, fx.data_ptr(), Nx, None if symmetric else Dy,
                            None if symmetric else fy.data_ptr(), Ny, tiles, symmetric, K.data_ptr(), Nx, stream)
            launches += 1
    launches += sx.launches + (sy.launches if sy is not None else 0)
    prof["gram_ms"] = ms
    prof["Dx"], prof["Dy"] = Dx, Dy
    prof["launches"] = launches
    if size > 1:
        K = reduce_sum_to_root(comm, K)
    out = None
    if rank == 0:
        out = K if return_device else K.cpu().numpy()
    torch.cuda.synchronize()
    prof["total_s"] = time.perf_counter() - t_all
    prof["no_converge"] = 0   # a sweep-limit hit raises above (on every rank)
    return out, prof
