"""Write a synthetic stand-in for ``datasets/elliptic_preproc.csv`` (the Kaggle Elliptic data set is
not available offline).  Same shape as the file elliptic_preproc.py:22-26 of the reference writes:
columns ``Unnamed: 0`` (the pandas index the reference keeps, which becomes feature 0),
``Class`` (0 = illicit, 1 = licit) and ``Feature 1`` .. ``Feature 165``; heavy-tailed values with a
class-dependent shift so that the SVM step is not degenerate.

    python make_synthetic_dataset.py [n_illicit=600] [n_licit=2400] [seed=0] [out=datasets/elliptic_synth.csv]
"""
import pathlib
import sys

import numpy as np
import pandas as pd


def make(n_illicit=600, n_licit=2400, seed=0, n_features=165):
    rng = np.random.default_rng(seed)
    n = n_illicit + n_licit
    cls = np.concatenate([np.zeros(n_illicit, dtype=int), np.ones(n_licit, dtype=int)])
    x = rng.standard_t(3, size=(n, n_features))
    shift = rng.normal(0.0, 0.6, size=n_features)
    x += np.where(cls[:, None] == 0, shift[None, :], -0.25 * shift[None, :])
    perm = rng.permutation(n)
    df = pd.DataFrame(x[perm], columns=[f"Feature {i + 1}" for i in range(n_features)])
    df.insert(0, "Class", cls[perm])
    df.insert(0, "Unnamed: 0", np.arange(n))
    return df


if __name__ == "__main__":
    a = sys.argv[1:]
    n_ill = int(a[0]) if len(a) > 0 else 600
    n_lic = int(a[1]) if len(a) > 1 else 2400
    seed = int(a[2]) if len(a) > 2 else 0
    out = pathlib.Path(a[3] if len(a) > 3 else "datasets/elliptic_synth.csv")
    out.parent.mkdir(parents=True, exist_ok=True)
    make(n_ill, n_lic, seed).to_csv(out, index=False)
    print(f"wrote {out}")
