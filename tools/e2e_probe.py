import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from bench import build_inputs
from nls_b200.native import nls
w = build_inputs("c2")
P = torch.from_numpy(np.ascontiguousarray(w["pumping"][0])).pin_memory().numpy()
u = torch.from_numpy(np.ascontiguousarray(w["u0"][0])).pin_memory().numpy()
c = w["coeffs"][0]
for iters in (0, 1, 32, 64, 640, 5000, 5000):
    nls.solve_nls_2d(w["dt"], w["dx"], 5, iters, P, c, u)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        nls.solve_nls_2d(w["dt"], w["dx"], 5, iters, P, c, u)
    dt = (time.perf_counter() - t0) / 3
    print("iters %5d: %.3f ms per call" % (iters, dt * 1e3))
