"""Generates pbrt_v3_b200/lib/sobol_matrices32.bin: the generator matrices of the first 1024 Sobol' dimensions as 52 columns
of 32-bit fractions each (column k = the direction number of index bit k) - the table the reference keeps as SobolMatrices32
(src/core/sobolmatrices.cpp, produced by L. Gruenschloss' generator from the Joe-Kuo direction numbers "new-joe-kuo-6.21201").

Nothing is read from the reference: the direction numbers come from scipy's copy of the Joe-Kuo set
(scipy.stats._sobol._initialize_v), computed at 52 bits and truncated to the upper 32.  The other two tables of the
reference (VdCSobolMatrices / VdCSobolMatricesInv, used by SobolIntervalToIndex) follow from dimensions 0 and 1 by linear
algebra over GF(2) and are derived inside the library (pb2_cuda.cu, sobolIntervalTables).  tests/test_oracle.py compares
sample values with ones recorded from the compiled reference's SobolSampler.

    python tools/make_sobol_tables.py            (run by __graft_entry__.build())
"""
import os
import sys

import numpy as np

N_DIMS, N_COLS = 1024, 52
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pbrt_v3_b200", "lib", "sobol_matrices32.bin")


def matrices32():
    from scipy.stats import _sobol
    v = np.zeros((N_DIMS, N_COLS), dtype=np.uint64)
    _sobol._initialize_v(v, dim=N_DIMS, bits=N_COLS)
    # dimension 0 is the van der Corput sequence: column k holds bit (51 - k) alone
    assert all(int(v[0, k]) == 1 << (N_COLS - 1 - k) for k in range(N_COLS))
    return (v >> np.uint64(N_COLS - 32)).astype("<u4")


def main():
    m = matrices32()
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "wb") as f:
        f.write(m.tobytes())
    print("wrote", OUT, m.shape, "checksum", int(m.astype(np.uint64).sum()))


if __name__ == "__main__":
    sys.exit(main())
