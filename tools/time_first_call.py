"""Where does the first-call latency go?  (context creation vs module load vs allocation)"""
import time, os, sys, ctypes as C
t0 = time.perf_counter()
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
t1 = time.perf_counter()
from kvxopt_b200 import _lib as L, cholmod
t2 = time.perf_counter()
print("import numpy/scipy %.2f s, import kvxopt_b200 %.2f s" % (t1 - t0, t2 - t1), flush=True)
t = time.perf_counter(); nd = L.device_count(); print("device_count=%d %.3f s" % (nd, time.perf_counter() - t), flush=True)
t = time.perf_counter(); L.fn["b200s_set_device"](0); print("set_device %.3f s" % (time.perf_counter() - t), flush=True)
n = 2000
A = sp.diags([-np.ones(n - 1), 4 * np.ones(n)], [-1, 0]).tocsc(); A.sort_indices()
t = time.perf_counter(); F = cholmod.symbolic(A); print("symbolic %.3f s" % (time.perf_counter() - t), flush=True)
for i in range(3):
    t = time.perf_counter(); cholmod.numeric(A, F); print("numeric call %d: %.4f s" % (i, time.perf_counter() - t), flush=True)
B = np.asfortranarray(np.ones((n, 1)))
for i in range(3):
    t = time.perf_counter(); cholmod.solve(F, B); print("solve call %d: %.4f s" % (i, time.perf_counter() - t), flush=True)
