import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, scipy.sparse as sp
import feastsolver_jl_b200 as fs
from oracle import feast_oracle as fo
n = 1000
A = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(n, n), format="csc")
X = (np.random.default_rng(3).standard_normal((n, 60)) + 1j * np.random.default_rng(4).standard_normal((n, 60))) / np.sqrt(2)
ct_o = fo.circular_contour_trapezoidal(0.05 + 0j, 0.05, 16); ct_g = fs.circular_contour_trapezoidal(0.05 + 0j, 0.05, 16)
print("oracle", fo.contour_estimate_eig(A, ct_o, X=X.copy()), "gpu", fs.contour_estimate_eig(A, ct_g, X=X.copy()))
exact = np.sum(2 - 2 * np.cos(np.arange(1, n + 1) * np.pi / (n + 1)) <= 0.1); print("exact", exact)
