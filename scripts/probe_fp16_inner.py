"""CPU experiment (oracle only): does the outer FGMRES notice when the INNER A_uu solve sees an fp16-rounded copy
of the matrix (per-row scaling, optional row-sum compensation on the diagonal block)? dt is chosen so that the
mass/stiffness ratio h^2 / (dt (mu + gamma)) equals that of config 3 (h = 1/128, dt = 1e-2)."""
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from util import cavity_prm, make_oracle

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
dt = (1.0 / n) ** 2 / ((1.0 / 128) ** 2 / 1e-2)


def fp16_copy(A, compensate, dim=3, dtype=np.float16):
    A = A.tocsr().copy()
    rowmax = np.maximum(abs(A).max(axis=1).toarray().ravel(), 1e-300)
    D = sp.diags(1.0 / rowmax)
    Ah = (D @ A).tocsr()
    Ah.data = Ah.data.astype(dtype).astype(np.float64)
    Ah = (sp.diags(rowmax) @ Ah).tocsr()
    if compensate:
        # per scalar row r and column component c: put the rounding loss of the row back on the diagonal block
        E = (A - Ah).tocsr()
        nn = A.shape[0] // dim
        for c in range(dim):
            sel = sp.csr_matrix((np.ones(nn), (np.arange(nn) * dim + c, np.arange(nn))), shape=(A.shape[1], nn))
            loss = np.asarray((E @ sel).sum(axis=1)).ravel()  # sum over column nodes of component c
            rows = np.arange(A.shape[0])
            cols = (rows // dim) * dim + c
            Ah = Ah + sp.csr_matrix((loss, (rows, cols)), shape=A.shape)
    return Ah.tocsr()


for label, filt in [("exact", None), ("fp16", lambda A: fp16_copy(A, False)), ("fp16+rowsum", lambda A: fp16_copy(A, True)),
                    ("bf16-like(8 bit)", lambda A: fp16_copy(A, False, dtype=np.float16)), ]:
    o = make_oracle(cavity_prm(3, dt=dt), (n,) * 3, (0, 0, 0), (1, 1, 1), a_inv=("bicgstab", 1e-1, 2000))
    o.a_inv_filter = filt
    t0 = time.perf_counter()
    for k in range(2):
        o.run_one_step(k == 0)
    print(label, "dt", dt, "time", round(time.perf_counter() - t0, 1))
    for h in o.history:
        print("   ", h)
