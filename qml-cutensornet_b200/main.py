"""Experiment driver with the reference's command line (reference main.py:75-93, README.md:73-84):

    python main.py <backend> <num_features> <layers> <gamma> <distance> <n_illicit> <n_licit> <data_seed> <data_file>

``<backend>`` is GPU or CPU and selects ``gpu_backend`` / ``cpu_backend`` exactly like the reference
(main.py:119-122); both are served by libqkmps.so.  Outputs keep the reference's names:
``kernels/{train,test}_<info>.npy``, ``<info>.json`` profiling files and ``data/{train,test}_<info>.npy``
with the SVC sweep (main.py:161-162,175,189,235-236).  One process per GPU: run under ``torchrun`` for
more than one GPU; mpi4py is not required.
"""
import pathlib
import sys
import time

import numpy as np

USAGE = ("\nCall script as 'python main.py <backend> <num_features> <layers> <gamma> <distance> <n_illicit> "
         "<n_licit> <data_seed> <data_file>'.\nThe value of <backend> must be either GPU or CPU.")
REG_SWEEP = [4, 3.5, 3, 2.5, 2, 1.5, 1, 0.5, 0.1, 0.05, 0.01]


def load_and_scale(data_file, n_illicit, n_licit, seed, num_features):
    """Sample, split 80/20 and scale like the reference (main.py:47-66,126-143): quantile -> standard ->
    min-max to [0, 2], fitted on the training split; keep the first ``num_features`` columns."""
    import pandas as pd
    from sklearn.model_selection import train_test_split
    from sklearn.preprocessing import MinMaxScaler, QuantileTransformer, StandardScaler
    df = pd.read_csv(pathlib.Path("datasets") / data_file)
    picked = pd.concat([df[df["Class"] == 0].sample(n_illicit, random_state=seed * 20 + 2),
                        df[df["Class"] == 1].sample(n_licit, random_state=seed * 46 + 9)], axis=0)
    tr, te = train_test_split(picked, stratify=picked["Class"], test_size=0.2, random_state=seed * 26 + 19)
    y_tr, y_te = np.array(tr.pop("Class"), dtype=int), np.array(te.pop("Class"), dtype=int)
    x_tr, x_te = np.array(tr), np.array(te)
    for scaler in (QuantileTransformer(output_distribution="normal", n_quantiles=min(1000, len(x_tr))),
                   StandardScaler(), MinMaxScaler((0, 2))):
        x_tr = scaler.fit_transform(x_tr)
        x_te = scaler.transform(x_te)
    return x_tr[:, :num_features], y_tr, x_te[:, :num_features], y_te


def svc_sweep(k_fit, y_fit, k_eval, y_eval):
    from sklearn.metrics import accuracy_score, precision_score, recall_score, roc_auc_score
    from sklearn.svm import SVC
    rows = []
    for c in REG_SWEEP:
        svc = SVC(kernel="precomputed", C=c, tol=1e-3)
        svc.fit(k_fit, y_fit)
        pred = svc.predict(k_eval)
        rows.append([c, accuracy_score(y_eval, pred), precision_score(y_eval, pred, zero_division=0),
                     recall_score(y_eval, pred), roc_auc_score(y_eval, pred)])
    return rows


def main(argv):
    if len(argv) <= 9:
        raise ValueError(USAGE)
    backend = str(argv[1])
    num_features, reps, gamma, distance = int(argv[2]), int(argv[3]), float(argv[4]), int(argv[5])
    n_illicit, n_licit, data_seed, data_file = int(argv[6]), int(argv[7]), int(argv[8]), str(argv[9])
    truncation_error = 1e-16                                      # main.py:73
    if backend == "GPU":
        from gpu_backend.kernel_state_ansatz import KernelStateAnsatz, build_kernel_matrix
    elif backend == "CPU":
        from cpu_backend.kernel_state_ansatz import KernelStateAnsatz, build_kernel_matrix
    else:
        raise ValueError(USAGE)
    import os
    from qkmps.engine import SingleComm
    from qkmps.synth import entanglement_graph
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        from qkmps.comm import init_from_env
        comm = init_from_env()
    else:
        comm = SingleComm()
    rank, root = comm.Get_rank(), 0

    emap = entanglement_graph(num_features, distance)
    x_tr, y_tr, x_te, y_te = load_and_scale(data_file, n_illicit, n_licit, data_seed, num_features)
    if rank == root:
        print(f"backend {backend}: {num_features} features, {reps} layers, gamma {gamma}, distance {distance}, "
              f"{len(x_tr)} train / {len(x_te)} test points, {comm.Get_size()} process(es)")
    pathlib.Path("kernels").mkdir(exist_ok=True)
    pathlib.Path("data").mkdir(exist_ok=True)
    ansatz = KernelStateAnsatz(num_qubits=num_features, reps=reps, gamma=gamma, entanglement_map=emap,
                               hadamard_init=True)
    tag = (f"Nf{num_features}_r{reps}_g{gamma}_p0.0_nn{distance}_mslinear_Ntr{n_illicit}_s{data_seed}_"
           f"{data_file.split('.')[0]}")
    train_info, test_info = "train_" + tag, "test_" + tag

    t0 = time.perf_counter()
    k_train = build_kernel_matrix(comm, ansatz, X=x_tr, info_file=train_info, truncation_error=truncation_error)
    t1 = time.perf_counter()
    k_test = build_kernel_matrix(comm, ansatz, X=x_tr, Y=x_te, info_file=test_info, truncation_error=truncation_error)
    t2 = time.perf_counter()
    if rank != root:
        return None
    print(f"Built kernel matrix on training set. Time: {round(t1 - t0, 3)} seconds")
    print(f"Built kernel matrix on test set. Time: {round(t2 - t1, 3)} seconds")
    np.save(f"kernels/{train_info}.npy", k_train)
    np.save(f"kernels/{test_info}.npy", k_test)
    test_results = svc_sweep(k_train, y_tr, k_test, y_te)
    train_results = svc_sweep(k_train, y_tr, k_train, y_tr)
    np.save(f"data/{train_info}.npy", train_results)
    np.save(f"data/{test_info}.npy", test_results)
    best = max(test_results, key=lambda r: r[4])
    print(f"best test AUC {best[4]:.3f} (accuracy {best[1]:.3f}) at C = {best[0]}")
    return dict(k_train=k_train, k_test=k_test, x_train=x_tr, x_test=x_te, test_results=test_results)


if __name__ == "__main__":
    main(sys.argv)
