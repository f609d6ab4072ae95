"""times the one-shot C-ABI call bssm_b200_psi_filter with pinned host outputs (the e2e leg of bench.py)"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bssm_b200
import data
from bssm_b200._lib import FilterOut, c_double_p

N, n = 65536, 1000
model = data.svm_model(n)
alpha_h = torch.empty((n + 1) * N, dtype=torch.float64).pin_memory()
w_h = torch.empty((n + 1) * N, dtype=torch.float64).pin_memory()
out = FilterOut()
out.alpha = C.cast(alpha_h.data_ptr(), c_double_p)
out.weights = C.cast(w_h.data_ptr(), c_double_p)
ms = model.c_struct()
L = bssm_b200.lib()
for i in range(5):
    t0 = time.perf_counter()
    rc = L.bssm_b200_psi_filter(C.byref(ms), N, 1, C.byref(out))
    dt = time.perf_counter() - t0
    print("call %d: %.1f ms (device %.1f ms) rc=%d -> %.3e p-s/s" % (i, dt * 1e3, out.device_ms, rc, N * n / dt))
