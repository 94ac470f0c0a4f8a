import sys, os, time, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from conftest import load_matrix, lower_ccs
from kvxopt_b200 import cholmod
Al = lower_ccs(load_matrix("bcsstk24")); n = Al.shape[0]
B = np.asfortranarray(np.ones((n, 1)))
cholmod.linsolve(Al, B.copy(order="F"))
os.environ["B200S_DEBUG"] = "1"
t = time.perf_counter(); cholmod.linsolve(Al, B.copy(order="F")); print("linsolve %.2f ms" % ((time.perf_counter() - t) * 1e3))
t = time.perf_counter(); cholmod.linsolve(Al, B.copy(order="F")); print("linsolve %.2f ms" % ((time.perf_counter() - t) * 1e3))
