"""Small end-to-end pass over every kernel for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_small.py        (also --tool racecheck / synccheck)
Cholesky (small + tiled fronts, solves, sys modes, diag, getfactor, sparse right-hand sides, complex Hermitian, shared Schur
complements through virtual ranks), KLU (factor, early-column kernel, batch refactor wave kernel, dense block, solves), the KKT
solvers."""
import os, sys
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kvxopt_b200 import _lib as L, cholmod, klu
from bench import lap3d_lower
def lap(nx):
    Al = lap3d_lower(nx); perm = np.zeros(Al.shape[0], np.int64); L.fn["b200s_grid_nd_perm"](nx, nx, nx, 64, L.ptr_i64(perm)); return Al, perm
for nx in (6, 14):
    Al, perm = lap(nx); n = Al.shape[0]
    F = cholmod.symbolic(Al, p=perm); cholmod.numeric(Al, F)
    for s in (0, 4, 5, 7, 8):
        X = np.ones((n, 3), order="F"); cholmod.solve(F, X, sys=s)
    cholmod.diag(F); cholmod.getfactor(F)
    print("chol", nx, cholmod.factor_info(F)["max_front_rows"], "ok", flush=True)
rng = np.random.default_rng(0)
M = sp.random(700, 700, density=0.01, random_state=rng, format="csc"); A = (M + M.T + sp.identity(700) * 20).tocsc()
Al = sp.tril(A).tocsc(); Al.sort_indices(); F = cholmod.symbolic(Al); cholmod.numeric(Al, F); X = np.ones((700, 2), order="F"); cholmod.solve(F, X)
print("chol rand ok", flush=True)
z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "bp_800.npz"))
K = sp.csc_matrix((z["values"], z["rowind"].astype(np.int64), z["colptr"]), shape=(int(z["n"]),) * 2)
Fs = klu.symbolic(K); Fn = klu.numeric(K, Fs)
vals = K.data[None, :] * (1 + 1e-3 * rng.uniform(-1, 1, size=(40, K.nnz)))
klu.refactor_batch(Fn, vals); B = rng.standard_normal((40, 2, K.shape[0])); klu.solve_batch(Fn, B); klu.solve_batch(Fn, B, trans="T")
klu.get_numeric(K, Fs, Fn)
z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ACTIVSg2000.npz"))
K = sp.csc_matrix((z["values"], z["rowind"].astype(np.int64), z["colptr"]), shape=(int(z["n"]),) * 2)
Fs = klu.symbolic(K); Fn = klu.numeric(K, Fs)
vals = K.data[None, :] * (1 + 1e-3 * rng.uniform(-1, 1, size=(33, K.nnz)))
klu.refactor_batch(Fn, vals); B = rng.standard_normal((33, 1, K.shape[0])); klu.solve_batch(Fn, B)
print("klu ok", flush=True)
# streaming refactor, LDL' semantics, device KKT solver
klu.refactor_batch_begin(Fn, vals); klu.refactor_batch_begin(Fn, vals); klu.refactor_batch_end(Fn); klu.refactor_batch_end(Fn)
cholmod.options["supernodal"] = 0
Al, perm = lap(8); F = cholmod.symbolic(Al, p=perm); cholmod.numeric(Al, F)
for s in (2, 3, 4, 5, 6):
    X = np.ones((Al.shape[0], 2), order="F"); cholmod.solve(F, X, sys=s)
cholmod.getfactor(F)
del cholmod.options["supernodal"]
from kvxopt_b200 import kkt
nk, mk, pk = 60, 150, 5
Gk = sp.vstack([sp.identity(nk), sp.random(mk - nk, nk, density=0.1, random_state=rng)]).tocsc()
Ak = (sp.random(pk, nk, density=0.3, random_state=rng) + sp.csc_matrix((np.ones(pk), (np.arange(pk), np.arange(pk))), shape=(pk, nk))).tocsc()
Mk = sp.random(nk, nk, density=0.1, random_state=rng); Hk = (Mk @ Mk.T + 0.1 * sp.identity(nk)).tocsc()
fac = kkt.chol2(Gk, {"l": mk, "q": [], "s": []}, Ak)
for rep in range(2):
    sol = fac({"di": 1.0 / rng.uniform(0.3, 3.0, mk)}, sp.tril(Hk).tocsc())
    x, y, zz = rng.standard_normal(nk), rng.standard_normal(pk), rng.standard_normal(mk); sol(x, y, zz)
print("streaming + ldl + kkt ok", flush=True)

# early-column kernel on a small pattern (threshold lowered), sparse right-hand sides, complex Hermitian matrices, dense 'chol'
os.environ["B200S_KLU_EARLY_MINW"] = "4"
z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "bp_800.npz"))
K = sp.csc_matrix((z["values"], z["rowind"].astype(np.int64), z["colptr"]), shape=(int(z["n"]),) * 2)
Fs = klu.symbolic(K); Fn = klu.numeric(K, Fs)
vals = K.data[None, :] * (1 + 1e-3 * rng.uniform(-1, 1, size=(40, K.nnz)))
assert not klu.refactor_batch(Fn, vals).any()
del os.environ["B200S_KLU_EARLY_MINW"]
Al, perm = lap(10); n = Al.shape[0]
F = cholmod.symbolic(Al, p=perm); cholmod.numeric(Al, F)
Bs = sp.random(n, 3, density=0.01, random_state=rng, format="csc"); Bs.sort_indices()
for s_ in (0, 4, 7):
    cholmod.spsolve(F, Bs, sys=s_)
Mz = sp.random(300, 300, density=0.02, random_state=rng, format="csc")
Az = (Mz + 1j * sp.random(300, 300, density=0.02, random_state=rng, format="csc")).tocsc(); Az = (Az + Az.getH()).tocsc()
Az = (Az + sp.diags(np.asarray(abs(Az).sum(axis=1)).ravel() + 1.0)).tocsc()
Azl = sp.tril(Az).tocsc(); Azl.sort_indices()
Xz = np.asfortranarray(np.ones((300, 2), dtype=complex)); cholmod.linsolve(Azl, Xz)
Fz = cholmod.symbolic(Azl); cholmod.numeric(Azl, Fz); cholmod.diag(Fz); cholmod.getfactor(Fz)
fc = kkt.chol(Gk, {"l": mk, "q": [], "s": []}, Ak)
solc = fc({"di": 1.0 / rng.uniform(0.3, 3.0, mk)}, Hk)
x, y, zz = rng.standard_normal(nk), rng.standard_normal(pk), rng.standard_normal(mk); solc(x, y, zz)
from kvxopt_b200 import dist as D
Al, perm = lap(16)
caps = [cholmod.symbolic(Al, p=perm) for _ in range(4)]
vr = D.VirtualRanks(caps, split=True, split_args=dict(min_flops=1e4, min_rows=64))
assert vr.factorize(Al.data) == Al.shape[0]
X = np.ones((Al.shape[0], 1), order="F"); cholmod.solve(caps[0], X)
print("early + spsolve + complex + kkt.chol + shared Schur complements (%d shared fronts) ok" % len(vr.splan), flush=True)
