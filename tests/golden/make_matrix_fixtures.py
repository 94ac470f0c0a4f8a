"""Convert the reference's test matrices (reference tests/*.mtx) into compact .npz CCS fixtures.

Run in the build container (the GPU box has no /root/reference):
    python tests/golden/make_matrix_fixtures.py
Parsing follows the reference's own reader (reference tests/test_sparse_solvers.py:36-68): the
MatrixMarket `symmetric` flag is IGNORED, i.e. bcsstk13/24 are loaded as their stored lower triangle.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

REF = "/root/reference/tests"
OUT = os.path.dirname(os.path.abspath(__file__))


def read_mtx_like_reference(path):
    I, J, V = [], [], []
    with open(path) as f:
        header_done = False
        for line in f:
            if line.startswith("%"):
                continue
            parts = line.split()
            if not header_done:
                m, n, nnz = int(parts[0]), int(parts[1]), int(parts[2])
                header_done = True
                continue
            I.append(int(parts[0]) - 1)
            J.append(int(parts[1]) - 1)
            V.append(float(parts[2]))
    A = sp.coo_matrix((V, (I, J)), shape=(m, n)).tocsc()   # duplicates summed, rows sorted (kvxopt spmatrix semantics)
    A.sort_indices()
    return A


for name in ("ACTIVSg2000", "bcsstk13", "bcsstk24", "bp_800"):
    A = read_mtx_like_reference(os.path.join(REF, name + ".mtx"))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), n=np.int64(A.shape[0]), colptr=A.indptr.astype(np.int64),
                        rowind=A.indices.astype(np.int32), values=A.data)
    print(name, A.shape, A.nnz, os.path.getsize(os.path.join(OUT, name + ".npz")))
