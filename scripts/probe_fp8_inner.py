"""CPU experiment (oracle only): inner A_uu solve on an fp8 (e4m3-like, 3 mantissa bits) copy of the matrix, with and
without row-sum compensation on the diagonal block. dt chosen so that the mass/stiffness ratio matches config 3."""
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from util import cavity_prm, make_oracle

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dt = (1.0 / n) ** 2 / ((1.0 / 128) ** 2 / 1e-2)


def round_mantissa(x, bits, min_exp):
    """round to `bits` explicit mantissa bits, flush below 2^min_exp to multiples of the subnormal step"""
    m, e = np.frexp(x)  # x = m * 2^e, 0.5 <= |m| < 1
    scale = 2.0 ** (bits + 1)
    out = np.ldexp(np.round(m * scale) / scale, e)
    step = 2.0 ** (min_exp - bits)
    small = np.abs(x) < 2.0 ** min_exp
    out[small] = np.round(x[small] / step) * step
    return out


def low_copy(A, bits, min_exp, compensate, dim=3):
    A = A.tocsr().copy()
    rowmax = np.maximum(abs(A).max(axis=1).toarray().ravel(), 1e-300)
    Ah = (sp.diags(1.0 / rowmax) @ A).tocsr()
    Ah.data = round_mantissa(Ah.data, bits, min_exp)
    Ah = (sp.diags(rowmax) @ Ah).tocsr()
    if compensate:
        E = (A - Ah).tocsr()
        nn = A.shape[0] // dim
        for c in range(dim):
            sel = sp.csr_matrix((np.ones(nn), (np.arange(nn) * dim + c, np.arange(nn))), shape=(A.shape[1], nn))
            loss = np.asarray((E @ sel).sum(axis=1)).ravel()
            rows = np.arange(A.shape[0])
            Ah = Ah + sp.csr_matrix((loss, (rows, (rows // dim) * dim + c)), shape=A.shape)
    return Ah.tocsr()


cases = [("exact", None), ("fp16 (10 bits)", lambda A: low_copy(A, 10, -14, False)),
         ("fp8 e4m3 + rowsum", lambda A: low_copy(A, 3, -6, True)),
         ("6 bits + rowsum", lambda A: low_copy(A, 6, -14, True)), ("6 bits", lambda A: low_copy(A, 6, -14, False)),
         ("fp8 e4m3 (3 bits)", lambda A: low_copy(A, 3, -6, False))]
if len(sys.argv) > 2:
    cases = [c for c in cases if c[0] in sys.argv[2:]]
for label, filt in cases:
    o = make_oracle(cavity_prm(3, dt=dt), (n,) * 3, (0, 0, 0), (1, 1, 1), a_inv=("bicgstab", 1e-1, 300))
    o.a_inv_filter = filt
    t0 = time.perf_counter()
    tot = {"a_inv": 0, "n": 0}
    for k in range(2):
        o.run_one_step(k == 0)
    its = [h[4] for h in o.history]
    print(f"{label:22s} FGMRES its per Newton it {its}  (sum {sum(its)})  inner its of the last solve {o.precond_stats['a_inv']} / {o.precond_stats['n']}"
          f"  time {time.perf_counter() - t0:.0f}s", flush=True)
