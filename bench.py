#!/usr/bin/env python
"""bench.py -- one JSON line for the quantum-kernel hot path on N B200s.

A "step" is one pass of the hot path over one batch: simulate every datapoint's ansatz circuit as
an MPS (stage 1), pack + exchange the states, and build the full train Gram matrix (stage 2).
Workload = BASELINE.json configs[2] (the config the 1/2/4/8-GPU metric is quoted on; it fits one
GPU): 50 qubits, 2 layers, distance 2, 1000 synthetic points, gamma = 1.0 (maximal bond dimension).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

`value`   Gram entries/s (unique entries N(N+1)/2 per step), inputs resident in HBM, CUDA events.
`e2e`     same metric through the reference-facing entry point gpu_backend.build_kernel_matrix with
          HOST numpy buffers (H2D of X and D2H of K inside the timed region).
`--impl reference` times the CPU restatement of the reference (oracle/, ITensors semantics) on the
          host cores; the reference's own backends (Julia/ITensors, pytket-cutensornet) do not
          install offline.
"""
import argparse
import json
import os
import pathlib
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "qml-cutensornet_b200"))
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (n_qubits, reps, distance, gamma, n_points)
    "c3": (50, 2, 2, 1.0, 1000),
    "c3_g0.1": (50, 2, 2, 0.1, 1000),
    "c2": (20, 2, 1, 0.5, 200),
    "c1": (10, 2, 1, 0.5, 40),
}
L2_FLUSH_BYTES = 256 << 20


# --------------------------------------------------------------------------------------- CPU arm
def _cpu_worker_init():
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass


def _cpu_sim_one(args):
    n, gates, x, cutoff = args
    from oracle.ansatz import bind_gate_list
    from oracle.mps_ref import simulate_mps
    t0 = time.perf_counter()
    m = simulate_mps(n, bind_gate_list(gates, x), cutoff, "itensors")
    return time.perf_counter() - t0, m.tensors


_G_TENSORS = None   # set in the parent before the fork pool is created (inherited, not pickled)


def _cpu_inner_chunk(pairs):
    from oracle.mps_ref import inner_prepared, prepare_for_inner
    tensors = _G_TENSORS
    prep = {}
    for (i, j) in pairs:
        for k in (i, j):
            if k not in prep:
                prep[k] = prepare_for_inner(tensors[k])
    out = []
    t0 = time.perf_counter()
    for (i, j) in pairs:
        out.append(abs(inner_prepared(prep[i][1], prep[j][0], prep[j][2])) ** 2)
    return time.perf_counter() - t0, len(pairs)


def cpu_reference_sample(workload, budget_s=14.0, cores=None):
    """Time the oracle port (ITensors semantics) on the host cores on a bounded sample of the workload
    and extrapolate to the whole job: est = N*t_circ/P + pairs*t_inner/P."""
    import multiprocessing as mp
    from oracle.ansatz import ansatz_gate_list, entanglement_graph
    from oracle.synth import synthetic_features
    n, r, d, g, N = WORKLOADS[workload]
    P = cores or len(os.sched_getaffinity(0))
    gates = ansatz_gate_list(n, r, g, entanglement_graph(n, d))
    X = synthetic_features(N, n, 0)
    t0 = time.perf_counter()
    tc, _ = _cpu_sim_one((n, gates, X[0], 1e-16))
    n_circ = int(min(N, max(P, (0.6 * budget_s) * P / max(tc, 1e-4))))
    global _G_TENSORS
    ctx = mp.get_context("fork")
    with ctx.Pool(P, initializer=_cpu_worker_init) as pool:
        w0 = time.perf_counter()
        res = pool.map(_cpu_sim_one, [(n, gates, X[i], 1e-16) for i in range(n_circ)], chunksize=max(1, n_circ // (4 * P)))
        wall_c = time.perf_counter() - w0
    t_circ = [t for t, _ in res]
    _G_TENSORS = [t for _, t in res]
    # inner products on the simulated sample
    rng = np.random.default_rng(1)
    tp, _ = _cpu_inner_chunk([(0, min(1, n_circ - 1))] * 8)
    tp /= 8
    n_pairs = int(max(P, (0.3 * budget_s) * P / max(tp, 1e-6)))
    idx = rng.integers(0, n_circ, size=(n_pairs, 2))
    chunks = [[(int(a), int(b)) for a, b in c] for c in np.array_split(idx, 4 * P) if len(c)]
    with ctx.Pool(P, initializer=_cpu_worker_init) as pool:
        w0 = time.perf_counter()
        res = pool.map(_cpu_inner_chunk, chunks)
        wall_p = time.perf_counter() - w0
    _G_TENSORS = None
    pairs_total = N * (N + 1) // 2
    per_circ_wall = wall_c / n_circ          # already divided by P through the pool
    per_pair_wall = wall_p / n_pairs
    est = N * per_circ_wall + pairs_total * per_pair_wall
    return {
        "value": pairs_total / est, "unit": "entries/s", "cores": P, "kind": "port",
        "sample": f"{n_circ} of {N} circuits + {n_pairs} of {pairs_total} inner products on {P} processes "
                  f"(BLAS 1 thread each), extrapolated to the whole job",
        "circuits_per_s": 1.0 / per_circ_wall, "inner_products_per_s": 1.0 / per_pair_wall,
        "median_s_per_circuit_1core": float(np.median(t_circ)),
        "median_s_per_inner_1core": float(np.median([t / c for t, c in res])),
        "sample_wall_s": time.perf_counter() - t0,
    }


# --------------------------------------------------------------------------------------- helpers
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[5 + k].strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm),
                       sm_mhz_max_seen=float(max(sm)))
        return out


def overlap_flops(chi, symmetric=True):
    """Algorithmic FLOPs of stage 2: per pair (y, x) and site k
    8*(2*cy[k-1]*cx[k-1]*cx[k] + 2*cy[k-1]*cy[k]*cx[k]) with the actual bond dimensions."""
    chi = np.asarray(chi, dtype=np.float64)
    N, nb = chi.shape
    tri = np.tril(np.ones((N, N))) if symmetric else np.ones((N, N))
    total = 0.0
    for k in range(1, nb):
        a = chi[:, k - 1]
        t1 = np.outer(a, chi[:, k - 1] * chi[:, k])
        t2 = np.outer(a * chi[:, k], chi[:, k])
        total += 16.0 * float(((t1 + t2) * tri).sum())
    return total


def sim_bytes(chi, ops):
    """Algorithmic HBM bytes of stage 1: every 2-site op reads and writes its two site
    tensors (final-state bond dimensions used for every visit of a bond), 1-site ops and gauge moves
    likewise for the tensors they touch."""
    chi = np.asarray(chi, dtype=np.float64)
    per_site = 16.0 * 2.0 * chi[:, :-1] * chi[:, 1:]        # bytes of site tensor s, per state
    total = 0.0
    for kind, site, *_ in ops:
        if kind <= 2:
            total += 2.0 * per_site[:, site].sum()
        elif kind <= 5:
            total += 2.0 * (per_site[:, site] + per_site[:, site + 1]).sum()
        elif kind == 16:
            total += 2.0 * (per_site[:, site] + per_site[:, site + 1]).sum()
        else:
            total += 2.0 * (per_site[:, site] + per_site[:, site - 1]).sum()
    return total + 8.0 * chi.shape[0] * (chi.shape[1] - 1)


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.load(open(p))
            return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------- arms
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, r, d, g, N = WORKLOADS[args.workload]
    vals = []
    last = None
    for _ in range(args.warmup):
        cpu_reference_sample(args.workload, budget_s=3.0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        last = cpu_reference_sample(args.workload, budget_s=args.cpu_budget)
        vals.append(last["value"])
    wall = time.perf_counter() - t0
    v = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "gram_entries_per_sec", "value": v, "unit": "entries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "complex128 (f64)",
        "data": "synthetic", "gpu_launches": 0,
        "config": {"workload": f"{args.workload}: {n} qubits, {r} layers, distance {d}, gamma {g}, {N} points, train Gram",
                   "note": "CPU restatement of the reference (ITensors semantics, numpy/LAPACK); the reference's Julia "
                           "and pytket-cutensornet backends do not install offline"},
        "cpu_baseline": {k: last[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": v, "unit": "entries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "circuits_per_s": last["circuits_per_s"],
    }
    line["cpu_baseline"]["value"] = v
    print(json.dumps(line))


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n, r, d, g, N = WORKLOADS[args.workload]

    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_base = cpu_reference_sample(args.workload, budget_s=args.cpu_budget)   # before CUDA init (fork pool)

    import torch
    import qkmps
    from qkmps.engine import SingleComm, build_gram
    from qkmps.synth import entanglement_graph, synthetic_features
    from gpu_backend.kernel_state_ansatz import KernelStateAnsatz, build_kernel_matrix
    qkmps.lib()
    if world > 1:
        from qkmps.comm import init_from_env
        comm = init_from_env("nccl")
        device = int(os.environ.get("LOCAL_RANK", "0"))
    else:
        comm, device = SingleComm(), 0
    torch.cuda.set_device(device)

    X = synthetic_features(N, n, 0)
    ans = KernelStateAnsatz(n, r, g, entanglement_graph(n, d))
    gates = ans.ansatz_circ.get_commands()
    plans = {}

    def plan_factory(cap, early_exit=False, parallel=False):
        key = (cap, bool(early_exit), bool(parallel))
        if key not in plans:
            plans[key] = qkmps.Plan(n, gates, qkmps.QK_TRUNC_PYTKET, 1e-16, cap,
                                    (qkmps.QK_PLAN_EARLY_EXIT if early_exit else 0) |
                                    (qkmps.QK_PLAN_PARALLEL if parallel else 0))
        return plans[key]

    from gpu_backend.kernel_state_ansatz import _initial_cap
    cap0 = args.chi if args.chi > 0 else _initial_cap(ans, 1e-16)
    X_dev = torch.from_numpy(X).to(f"cuda:{device}")
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=f"cuda:{device}")
    pairs = N * (N + 1) // 2

    def barrier():
        if world > 1:
            comm.Barrier()
        torch.cuda.synchronize()

    def step_device():
        return build_gram(comm, plan_factory, n, X_dev, None, chi_cap=cap0, device=device, return_device=True)

    def step_e2e():
        return build_kernel_matrix(comm, ans, X, truncation_error=1e-16, chi=cap0)

    for _ in range(max(args.warmup, 3)):
        step_device()
        step_e2e()

    # ---- `value`: inputs resident in HBM, CUDA events, max over ranks
    sampler = ClockSampler(device)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sim_ms, gram_ms, launches, prof = [], [], 0, None
    barrier()
    if rank == 0:
        sampler.start()
    for k in range(args.steps):
        flush.zero_()                      # L2 flush between timed iterations
        barrier()
        ev[k][0].record()
        _, prof = step_device()
        ev[k][1].record()
        sim_ms.append(prof["sim_ms_x"]); gram_ms.append(prof["gram_ms"]); launches += prof["launches"]
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    # ---- `e2e`: host buffers through the reference-facing entry point
    e2e_s = []
    for k in range(args.steps):
        flush.zero_()
        barrier()
        t0 = time.perf_counter()
        Kh = step_e2e()
        torch.cuda.synchronize()
        e2e_s.append(time.perf_counter() - t0)
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    tot_ms = float(np.sum(step_ms))
    e2e_tot = float(np.sum(e2e_s))
    if world > 1:
        t = torch.tensor([tot_ms, e2e_tot, float(np.mean(sim_ms)), float(np.mean(gram_ms))], dtype=torch.float64,
                         device=f"cuda:{device}")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        tot_ms, e2e_tot, sim_mean, gram_mean = [float(v) for v in t.tolist()]
        lt = torch.tensor([launches], dtype=torch.int64, device=f"cuda:{device}")
        torch.distributed.all_reduce(lt)
        launches = int(lt.item())
    else:
        sim_mean, gram_mean = float(np.mean(sim_ms)), float(np.mean(gram_ms))
    if rank != 0:
        return

    ms_per_step = tot_ms / args.steps
    value = pairs / (ms_per_step * 1e-3)
    e2e_value = pairs / (e2e_tot / args.steps)

    # ---- roofline of the dominant kernel (per launch, this rank's share)
    info = prof["info_x"]
    lo, hi = prof["shard"]
    hbm_peak, hbm_src = measured_peaks()
    try:
        dmma_peak = qkmps.dmma_peak(device, 20000)
    except Exception:
        dmma_peak = None
    # stage 2 flops of rank 0's tiles ~ total / world (row blocks are dealt evenly)
    chi_all = info["chi"]
    if world == 1:
        f2 = overlap_flops(chi_all)
    else:
        f2 = overlap_flops(chi_all) * (N / max(hi - lo, 1)) ** 2 / world   # estimate from rank 0's shard
    b1 = sim_bytes(chi_all, prof["plan_obj"].ops())
    stage2 = {"bound": "tensor", "achieved": f2 / (gram_mean * 1e-3) / 1e12, "peak": dmma_peak, "unit": "TFLOP/s",
              "frac": (f2 / (gram_mean * 1e-3) / 1e12 / dmma_peak) if dmma_peak else None, "traffic": None,
              "kernel": "qk_gram_dmma_kernel", "ms": gram_mean,
              "peak_source": "DMMA m8n8k4 FP64 microbenchmark run in this process (MEASURED_PEAKS.json has no FP64 figure; "
                             "nominal 37-40 TFLOP/s)",
              "algorithmic_flops_per_launch": f2}
    stage1 = {"bound": "hbm", "achieved": b1 / (sim_mean * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
              "frac": b1 / (sim_mean * 1e-3) / 1e9 / hbm_peak, "traffic": None, "kernel": "qk_sim_kernel", "ms": sim_mean,
              "peak_source": hbm_src, "algorithmic_bytes_per_launch": b1,
              "note": "stage 1 is bound by FP64 Jacobi-SVD arithmetic and barrier latency, not HBM: the algorithmic "
                      "byte model of SURVEY.md 8(d) is reported as asked, see DESIGN.md"}
    # DRAM traffic per launch from the committed ncu captures (valid for the default workload on one GPU)
    try:
        tr = json.load(open(ROOT / "profiles" / "r01_dram_traffic.json"))
        if world == 1 and args.workload == "c3" and args.points == 0:
            stage1["traffic"] = tr["qk_sim_kernel"]["dram_bytes"]
            stage2["traffic"] = tr["qk_gram_dmma_kernel"]["dram_bytes"]
            stage1["traffic_source"] = tr["qk_sim_kernel"]["capture"]
            stage2["traffic_source"] = tr["qk_gram_dmma_kernel"]["capture"]
    except Exception:
        pass
    dominant = stage2 if gram_mean >= sim_mean else stage1
    roofline = {k: dominant[k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic")}
    roofline["kernel"] = dominant["kernel"]
    roofline["peak_source"] = dominant["peak_source"]

    line = {
        "metric": "gram_entries_per_sec", "value": value, "unit": "entries/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "complex128 (f64)", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {n} qubits, {r} layers, distance {d}, gamma {g}, {N} points, full train "
                               f"Gram ({pairs} unique entries)", "truncation_error": 1e-16, "trunc_rule": "pytket",
                   "l2": f"flushed with a {L2_FLUSH_BYTES >> 20} MiB write between timed iterations; packed states "
                         f"({prof['frag_bytes_per_state'][0] * N >> 20} MiB) exceed the 126 MB L2",
                   "parallelism": f"datapoints sharded over {world} GPU(s); all-gather of packed states; row blocks of K "
                                  "dealt to ranks",
                   "stage1_schedule": prof.get("stage1_schedule", ""), "stage2_kernel": prof.get("gram_kernel", "")},
        "circuits_per_s": N / (ms_per_step * 1e-3),
        "stage_ms": {"simulate": sim_mean, "gram": gram_mean, "other": ms_per_step - sim_mean - gram_mean},
        "roofline": roofline, "stages": {"simulate": stage1, "gram": stage2},
        "e2e": {"value": e2e_value, "unit": "entries/s", "h2d_bytes_per_step": int(X.nbytes),
                "d2h_bytes_per_step": int(N * N * 8), "ms_per_step": 1e3 * e2e_tot / args.steps},
        "gpu_launches": launches, "clocks": clocks,
        "max_chi": int(chi_all.max()), "mean_max_chi": float(chi_all.max(axis=1).mean()),
    }
    if cpu_base is not None:
        line["cpu_baseline"] = {k: cpu_base[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line["cpu_baseline"]["circuits_per_s"] = cpu_base["circuits_per_s"]
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-budget", type=float, default=14.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--chi", type=int, default=0, help="first bond cap tried (0: the backend's own choice)")
    ap.add_argument("--points", type=int, default=0, help="override the number of datapoints (experiments only)")
    args = ap.parse_args()
    if args.points > 0:
        n, r, d, g, _ = WORKLOADS[args.workload]
        WORKLOADS[args.workload] = (n, r, d, g, args.points)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
