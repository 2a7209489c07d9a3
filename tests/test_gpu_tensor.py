"""GPU tests of the tcgen05 building blocks: the UMMA descriptor probe and the hi/lo-split
tensor-core GEMM against float64 numpy and against the fp32 CUDA-core GEMM."""
import ctypes
from ctypes import POINTER, c_float, c_int, c_int64, c_uint16

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _hooks(gpu_lib):
    L = gpu_lib.lib
    L.ffb_test_umma_probe.restype = c_int
    L.ffb_test_umma_probe.argtypes = [POINTER(c_uint16), POINTER(c_uint16), POINTER(c_float), c_int, c_int]
    L.ffb_test_umma_probe2.restype = c_int
    L.ffb_test_umma_probe2.argtypes = [POINTER(c_uint16), POINTER(c_uint16), POINTER(c_float), c_int, c_int, c_int]
    L.ffb_test_gemm.restype = c_int
    L.ffb_test_gemm.argtypes = [POINTER(c_float), POINTER(c_float), POINTER(c_float), POINTER(c_float), c_int64, c_int,
                                c_int, c_int, POINTER(c_float)]
    return L


@pytest.mark.parametrize("N,K", [(16, 16), (80, 32), (64, 256), (96, 256), (256, 64)])
def test_umma_probe_noswizzle(gpu_lib, N, K):
    L = _hooks(gpu_lib)
    rng = np.random.default_rng(N + K)
    A = rng.uniform(-1, 1, (128, K)).astype(np.float16)
    B = rng.uniform(-1, 1, (N, K)).astype(np.float16)
    D = np.zeros((128, N), np.float32)
    r = L.ffb_test_umma_probe(A.view(np.uint16).ctypes.data_as(POINTER(c_uint16)),
                              B.view(np.uint16).ctypes.data_as(POINTER(c_uint16)),
                              D.ctypes.data_as(POINTER(c_float)), N, K)
    assert r == 0
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    assert np.max(np.abs(D - ref)) < 1e-4 * max(1.0, K / 64)


@pytest.mark.parametrize("N,K", [(16, 16), (16, 256), (80, 64), (48, 256)])
def test_umma_probe_a_in_tmem(gpu_lib, N, K):
    """A operand resident in tensor memory (tcgen05.st, lane = row, column c = halfs 2c, 2c+1):
    the layout the recurrent kernel keeps its weight slice in."""
    L = _hooks(gpu_lib)
    rng = np.random.default_rng(3 * N + K)
    A = rng.uniform(-1, 1, (128, K)).astype(np.float16)
    B = rng.uniform(-1, 1, (N, K)).astype(np.float16)
    D = np.zeros((128, N), np.float32)
    r = L.ffb_test_umma_probe2(A.view(np.uint16).ctypes.data_as(POINTER(c_uint16)),
                               B.view(np.uint16).ctypes.data_as(POINTER(c_uint16)),
                               D.ctypes.data_as(POINTER(c_float)), N, K, 1)
    assert r == 0
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    assert np.max(np.abs(D - ref)) < 1e-4 * max(1.0, K / 64)


def gemm(L, A, W, b, mode):
    M, K = A.shape
    N = W.shape[0]
    C = np.zeros((M, N), np.float32)
    ms = c_float(0)
    r = L.ffb_test_gemm(A.ctypes.data_as(POINTER(c_float)), W.ctypes.data_as(POINTER(c_float)),
                        b.ctypes.data_as(POINTER(c_float)), C.ctypes.data_as(POINTER(c_float)), M, N, K, mode,
                        ctypes.byref(ms))
    assert r == 0
    return C, ms.value


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (1000, 768, 256), (333, 192, 64), (4097, 1024, 256),
                                   (700, 1536, 384), (129, 384, 128), (8300, 1536, 384), (65, 128, 320),
                                   (700, 2048, 512), (129, 1536, 512), (300, 192, 512)])   # K = 512: the K-split accumulators / the plain A-tile kernel
def test_gemm_tc_matches_fp64(gpu_lib, M, N, K):
    L = _hooks(gpu_lib)
    rng = np.random.default_rng(M + N + K)
    A = rng.uniform(-1, 1, (M, K)).astype(np.float32)
    W = (rng.uniform(-1, 1, (N, K)) * 3 / np.sqrt(K)).astype(np.float32)
    b = rng.uniform(-.3, .3, N).astype(np.float32)
    ref = A.astype(np.float64) @ W.astype(np.float64).T + b
    C_tc, _ = gemm(L, A, W, b, 0)
    C_f32, _ = gemm(L, A, W, b, 1)
    err_tc = np.max(np.abs(C_tc - ref))
    err_f32 = np.max(np.abs(C_f32 - ref))
    # the split-fp16 tensor path must be as accurate as a plain fp32 GEMM (same order of error)
    assert err_f32 < 5e-6
    # 22-bit operands + tensor-core fp32 accumulation: within a small factor of plain fp32
    assert err_tc < 2e-5 and err_tc < 6 * err_f32, (err_tc, err_f32)


def test_gemm_tc_speed_report(gpu_lib):
    L = _hooks(gpu_lib)
    rng = np.random.default_rng(0)
    M, N, K = 262144, 768, 256
    A = rng.uniform(-1, 1, (M, K)).astype(np.float32)
    W = (rng.uniform(-1, 1, (N, K)) * 3 / np.sqrt(K)).astype(np.float32)
    b = np.zeros(N, np.float32)
    C_tc, ms_tc = gemm(L, A, W, b, 0)
    C_f32, ms_f32 = gemm(L, A, W, b, 1)
    fl = 2.0 * M * N * K
    print(f"\nGEMM {M}x{N}x{K}: tcgen05(split+gemm) {ms_tc:.3f} ms = {fl/ms_tc/1e9:.1f} TFLOP/s algorithmic; "
          f"fp32 SIMT {ms_f32:.3f} ms = {fl/ms_f32/1e9:.1f} TFLOP/s")
    assert np.max(np.abs(C_tc - C_f32)) < 2e-5


@pytest.mark.parametrize("M,N,K,hop", [(300, 128, 320, 80), (1000, 384, 320, 80), (257, 256, 64, 16)])
def test_conv_gemm_im2col_view(gpu_lib, M, N, K, hop):
    """The tensor-core convolution reads its A operand through an OVERLAPPING-row tensor map (row r = K elements from
    element r * hop of the activation planes): compare with an explicit im2col in float64."""
    L = _hooks(gpu_lib)
    L.ffb_test_conv_gemm.restype = c_int
    L.ffb_test_conv_gemm.argtypes = [POINTER(c_float), ctypes.c_int64, ctypes.c_int64, POINTER(c_float), POINTER(c_float),
                                     POINTER(c_float), ctypes.c_int64, c_int, c_int]
    rng = np.random.default_rng(M + N)
    nx = (M - 1) * hop + K
    x = rng.uniform(-1, 1, nx).astype(np.float32)
    W = (rng.uniform(-1, 1, (N, K)) * 2 / np.sqrt(K)).astype(np.float32)
    b = rng.uniform(-.3, .3, N).astype(np.float32)
    out = np.zeros((M, N), np.float32)
    r = L.ffb_test_conv_gemm(x.ctypes.data_as(POINTER(c_float)), nx, hop, W.ctypes.data_as(POINTER(c_float)),
                             b.ctypes.data_as(POINTER(c_float)), out.ctypes.data_as(POINTER(c_float)), M, N, K)
    assert r == 0
    A = np.lib.stride_tricks.as_strided(x, shape=(M, K), strides=(4 * hop, 4)).astype(np.float64)
    z = A @ W.astype(np.float64).T + b
    ref = z / (1.0 + np.exp(-z))                                   # swish
    assert np.max(np.abs(out - ref)) < 2e-5
