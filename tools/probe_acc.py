import sys; sys.path.insert(0,'/root/repo')
import numpy as np, ctypes
from ctypes import POINTER, c_float, c_int, c_uint16
from flappie_b200.api import Library
L = Library.get().lib
L.ffb_test_umma_probe.restype = c_int
L.ffb_test_umma_probe.argtypes = [POINTER(c_uint16), POINTER(c_uint16), POINTER(c_float), c_int, c_int]
rng = np.random.default_rng(0)
for K in (16, 32, 64, 128, 256, 512):
    N = 64
    A = np.abs(rng.uniform(0.2, 1, (128, K))).astype(np.float16)   # all positive -> monotone growing accumulator
    B = np.abs(rng.uniform(0.2, 1, (N, K))).astype(np.float16)
    D = np.zeros((128, N), np.float32)
    assert L.ffb_test_umma_probe(A.view(np.uint16).ctypes.data_as(POINTER(c_uint16)), B.view(np.uint16).ctypes.data_as(POINTER(c_uint16)), D.ctypes.data_as(POINTER(c_float)), N, K) == 0
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    rel = (D - ref) / ref
    # fp32 sequential sum for comparison
    f32 = np.zeros((128, N), np.float32)
    for k in range(K):
        f32 += (A[:, k:k+1].astype(np.float32) * B[:, k].astype(np.float32)[None, :])
    relf = (f32 - ref) / ref
    print(f"K={K:4d} n_mma={K//16:3d}  TC rel err: mean {rel.mean():+.3e} max|.| {np.abs(rel).max():.3e}   fp32-seq: mean {relf.mean():+.3e} max {np.abs(relf).max():.3e}   (2^-24={2**-24:.2e})")
