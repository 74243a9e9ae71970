import sys, os, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import numpy as np
import cases, helpers
from structurepreservingiterativesolvers_b200 import solvers
g = np.load(os.path.join(ROOT, "tests", "golden", "reference_outputs.npz"))
name = "heat_tol7_jacobi"
out = {}
for fmt in ("sell", "pattern", "csr"):
    solvers.configure(spmv_format=fmt)
    x, info, dic, prob = helpers.run_product(name, orth="mgs", lookahead=False)
    X = g[name + "/X"]
    print(fmt, "steps", info["steps"], [("%.1e" % helpers.rel_diff(info["x"][j], X[j])) for j in range(1, min(6, len(X)))], "res", info["res"][:3])
    out[fmt] = [np.array(info["x"][j]) for j in range(len(X))]
solvers.configure(spmv_format="auto")
for j in range(1, 5):
    print(j, "sell vs pattern %.2e" % helpers.rel_diff(out["sell"][j], out["pattern"][j]), "sell vs csr %.2e" % helpers.rel_diff(out["sell"][j], out["csr"][j]))
