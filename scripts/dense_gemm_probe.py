import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import feastsolver_jl_b200 as fs
from feastsolver_jl_b200 import workloads as wl
n, m0 = int(sys.argv[1]), int(sys.argv[2])
A = wl.dense_nonhermitian(n, seed=1)
with fs.FeastContext() as ctx:
    ctx.set_operator(0, A); ctx.set_problem(0, 1, n)
    ctx.set_subspace(wl.rand_subspace(n, m0, seed=0))
    _, ms = ctx.apply_operator(0, which=0, download=False, reps=5)
    print(json.dumps({"zgemm": [n, m0, n], "ms": ms, "tflops": 8 * n * n * m0 / ms / 1e9}))
    if len(sys.argv) > 3:
        F = ctx.factorize([1.0, -0.3 - 0.7j]); ctx.factor_free(F)
