"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the
header declares, and refuses to compute without a GPU (no silent CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from flappie_b200 import api
from flappie_b200.model import KIND_GRU, KIND_LSTM, FlipflopModel, Mat, conv_mat_image, mat_image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "flappie_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?[A-Za-z_][\w\s\*]*?\b(\w+)\s*\([^;{]*\)\s*;", src, flags=re.M)
    return sorted(set(n for n in names if n not in ("defined",)))


def test_library_exports_every_declared_symbol(lib):
    declared = header_functions()
    assert len(declared) >= 30
    assert set(declared) == set(api.EXPORTS), set(declared) ^ set(api.EXPORTS)
    for name in declared:
        assert hasattr(lib.lib, name), name


def test_model_registry_names(lib):
    L = lib.lib
    # reference src/networks.c:21-83
    for name, val in (("r941_native", 0), ("r941_rna002", 1), ("r941_5mC", 2), ("r103_native", 3)):
        assert L.get_flappie_model_type(name.encode()) == val
        assert L.flappie_model_string(val).decode() == name
        assert len(L.flappie_model_description(val)) > 10
    assert L.get_flappie_model_type(b"r10C_pcr") == 0          # flappie-1.x alias
    assert L.get_flappie_model_type(b"nonsense") == 4          # FLAPPIE_MODEL_INVALID
    assert L.get_flappie_model_type(b"rle_r941_native") == 5
    for nparam, nbase in ((40, 4), (60, 5), (12, 2)):
        assert L.nbase_from_flipflop_nparam(nparam) == nbase


def test_matrix_ownership_roundtrip(lib):
    # results are callee-allocated `_Mat`s the caller frees (flappie_matrix.c:20-51,142)
    m = lib.lib.make_flappie_matrix(10, 7)
    assert m.contents.nr == 10 and m.contents.nrq == 3 and m.contents.stride == 12 and m.contents.nc == 7
    buf = np.ctypeslib.as_array(m.contents.data, shape=(7, 12))
    assert not buf.any()                                        # zero-initialised incl. padding
    assert not lib.lib.free_flappie_matrix(m)                   # returns NULL
    rows = np.arange(21, dtype=np.float32).reshape(3, 7)
    assert np.array_equal(lib.rows_from_mat(lib.mat_from_rows(rows)), rows)
    assert not lib.lib.make_flappie_matrix(0, 3)


def test_null_in_null_out(lib):
    # RETURN_NULL_IF semantics (flappie_stdlib.h:44)
    assert not lib.lib.transpost_crf_flipflop(None, True)
    assert not lib.lib.trace_from_posterior(None)
    assert np.isnan(lib.lib.decode_crf_flipflop(None, False, None, None))
    rt = api.RawTable(None, 0, 0, 0, None)
    assert not lib.lib.calculate_transitions(rt, 1.0, 0)
    assert lib.lib.ffb_basecall_batch(None, None) != 0
    assert not lib.lib.ffb_create(None, None)


def test_emit_bases_matches_oracle(lib, oracle):
    rng = np.random.default_rng(4)
    for nbase in (4, 5):
        for _ in range(20):
            T = int(rng.integers(2, 400))
            path = np.repeat(rng.integers(0, 2 * nbase, size=T), rng.integers(1, 4, size=T))[:T + 1].astype(np.int32)
            path = np.resize(path, T + 1)
            qpath = np.log(rng.uniform(1e-4, 1.0, size=T + 1)).astype(np.float32)
            assert lib.emit_bases(path, qpath, nbase) == oracle.emit_bases(path, qpath, nbase)
            b, q = lib.emit_bases(path, qpath, nbase, reverse=True)
            b2, q2 = oracle.emit_bases(path, qpath, nbase)
            assert b == b2[::-1] and q == q2[::-1]
    # quality clipping (util.h:285-305): p -> 1 saturates at Q50 = 'S'
    b, q = lib.emit_bases(np.array([0, 1, 1], np.int32), np.array([np.nan, 0.0, 0.0], np.float32), 4)
    assert b == "C" and q == "S"


def test_phred_table_reproduces_host_quality_chars(lib):
    """The device emission (emit.cu) only compares qpath against ffb_phred_table(): that lookup must give the very
    characters ffb_emit_bases computes with the host's expf / log1pf (reference src/util.h:285-305), everywhere --
    random values, a dense sweep, and the floats next to every step."""
    thr = lib.phred_table()
    assert 40 <= len(thr) <= 60 and np.all(np.diff(thr) > 0)
    rng = np.random.default_rng(9)
    xs = [rng.uniform(-12, 0.5, 200000).astype(np.float32), np.linspace(-3e-5, 1e-6, 50001, dtype=np.float32),
          np.array([-np.inf, -1000.0, -88.0, 0.0, 1.0, 80.0, np.inf, np.nan], np.float32)]
    for t in thr:                                       # neighbours of every step, a few ulps either side
        k = np.float32(t).view(np.int32)
        xs.append((k + np.arange(-4, 5, dtype=np.int32)).astype(np.int32).view(np.float32))
    x = np.concatenate(xs)
    look = 33 + np.searchsorted(thr, x, side="right")
    look[np.isnan(x)] = 33 + len(thr)
    path = (np.arange(len(x) + 2) % 2).astype(np.int32)              # nblock = len(x) + 1: a base at every pos in [1, nblock)
    qpath = np.concatenate([[0.0], x, [0.0]]).astype(np.float32)
    b, q = lib.emit_bases(path, qpath, 4)
    assert len(q) == len(x)
    got = np.frombuffer(q.encode("ascii"), np.uint8)
    assert np.array_equal(got, look.astype(np.uint8)), np.flatnonzero(got != look)[:10]
    assert got.max() == 83 and got.min() == 33                       # clipped at Q50 = 'S'


def test_mat_bundle_layout():
    # the `_Mat` images are exactly what the generated model headers define
    # (misc/taiyaki_flipflop5_guppy.py:38-99)
    W = np.arange(2 * 5 * 3, dtype=np.float32).reshape(2, 5, 3)   # nfilter 2, winlen 5, nf 3
    m, buf = conv_mat_image(W)
    assert (m.nr, m.nrq, m.nc, m.stride) == (4 * 5 - 4 + 3, 5, 2, 20)
    assert buf[1, 4 * 2 + 1] == W[1, 2, 1] and buf[0, 3] == 0.0   # 4th feature slot is padding
    m, buf = mat_image(np.ones((6, 10), np.float32))
    assert (m.nr, m.nrq, m.nc, m.stride) == (10, 3, 6, 12)
    for kind, n in ((KIND_GRU, 19), (KIND_LSTM, 23)):
        fm = FlipflopModel.synthetic(kind, 64, 4, seed=1)
        mats, keep = fm.to_mat_bundle()
        assert len(mats) == n
        assert fm.nparam == 40 and fm.nbase == 4
    fm = FlipflopModel.synthetic(KIND_GRU, 64, 5, seed=1)
    assert fm.nblock(3790) == 1895 and fm.nblock(10) == -1
    fm = FlipflopModel.synthetic(KIND_LSTM, 96, 4, seed=1)
    assert fm.nblock(3790) == 758 and fm.stride == 5


def test_no_silent_cpu_fallback(lib):
    if lib.device_count() > 0:
        pytest.skip("a GPU is present; the loud-failure path is exercised on CPU boxes")
    fm = FlipflopModel.synthetic(KIND_GRU, 64, 4, seed=1)
    with pytest.raises(api.FlappieB200Error):
        api.Model(fm)
    with pytest.raises(api.FlappieB200Error):
        lib.decode_crf_flipflop(np.zeros((4, 40), np.float32))
    # the raw C entry points report failure the reference's way: NULL / NAN
    tm = lib.mat_from_rows(np.zeros((4, 40), np.float32))
    assert not lib.lib.transpost_crf_flipflop(tm, True)
    path = (ctypes.c_int * 8)(); q = (ctypes.c_float * 8)()
    assert np.isnan(lib.lib.decode_crf_flipflop(tm, False, path, q))
    lib.lib.free_flappie_matrix(tm)
    assert "CUDA" in lib.last_error() or "device" in lib.last_error()


def test_product_never_imports_oracle():
    # a product path that routes through the oracle voids every parity claim
    pkg = os.path.join(ROOT, "flappie_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in src and "flappie_oracle" not in src and "libflappie_ref" not in src, f


def test_emit_runs_matches_oracle(lib, oracle):
    """ffb_emit_runs (host C++, runnie.c:279-310) against the oracle's loop on random state paths: no device needed."""
    rng = np.random.default_rng(5)
    for T in (0, 1, 2, 50, 3000):
        path = rng.integers(0, 8, size=T).astype(np.int32)
        if T > 10:
            path[:7] = 5                      # leading stay states before the first base: counted into nothing
        post = rng.uniform(0.5, 9.0, size=(max(T, 1), 40)).astype(np.float32)[:T]
        rle = np.ascontiguousarray(post[:, :8])
        b_g, sh_g, sc_g, dw_g = lib.emit_runs(path, rle) if T else ("", np.zeros(0), np.zeros(0), np.zeros(0))
        b_o, sh_o, sc_o, dw_o = oracle.emit_runs(path, post) if T else ("", np.zeros(0), np.zeros(0), np.zeros(0))
        assert b_g == b_o and np.array_equal(dw_g, dw_o) and np.array_equal(sh_g, sh_o) and np.array_equal(sc_g, sc_o)
        assert len(b_g) == int(np.sum(path < 4))


def test_device_only_entry_points_fail_loudly_without_gpu(lib):
    """New entry points of this round keep the rule: no device -> NULL / negative status, never a CPU path."""
    import ctypes
    if lib.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    L = lib.lib
    L.ffb_alloc_pinned.restype = ctypes.c_void_p; L.ffb_alloc_pinned.argtypes = [ctypes.c_size_t]
    assert not L.ffb_alloc_pinned(1 << 20)
    p = np.zeros((5, 40), np.float32)
    with pytest.raises(Exception):
        lib.decode_crf_runlength(p)
    with pytest.raises(Exception):
        lib.transpost_crf_runlength(p)


def test_round2_entry_points_without_gpu(lib, tmp_path):
    """ffb_model_load / lazy registration / the host-only helpers added in round 2: the host parts work without a device, the
    device parts refuse loudly (NULL + message), nothing computes on the CPU."""
    import ctypes
    L = lib.lib
    # host-only: change_positions (decode.c:66-79), nbase_from_crf_runlength_nparam (layers.c:1235), array_from_flappie_imatrix
    L.change_positions.restype = ctypes.c_size_t
    L.change_positions.argtypes = [ctypes.POINTER(ctypes.c_int), ctypes.c_size_t, ctypes.POINTER(ctypes.c_int)]
    path = np.array([3, 3, 1, 1, 1, 5, 2, 2], np.int32)
    ch = np.zeros(8, np.int32)
    n = L.change_positions(path.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), 8, ch.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    assert n == 3 and ch[:3].tolist() == [2, 5, 6]
    assert L.change_positions(None, 8, ch.ctypes.data_as(ctypes.POINTER(ctypes.c_int))) == 0          # RETURN_NULL_IF
    L.nbase_from_crf_runlength_nparam.restype = ctypes.c_size_t
    L.nbase_from_crf_runlength_nparam.argtypes = [ctypes.c_size_t]
    assert L.nbase_from_crf_runlength_nparam(40) == 4
    L.array_from_flappie_imatrix.restype = ctypes.POINTER(ctypes.c_int32)
    L.array_from_flappie_imatrix.argtypes = [ctypes.POINTER(api.IMat)]
    assert not L.array_from_flappie_imatrix(None)
    im = L.make_flappie_imatrix(3, 2)
    np.ctypeslib.as_array(im.contents.data, shape=(2, 4))[:] = [[1, 2, 3, 0], [4, 5, 6, 0]]
    dense = L.array_from_flappie_imatrix(im)
    assert np.ctypeslib.as_array(dense, shape=(6,)).tolist() == [1, 2, 3, 4, 5, 6]
    ctypes.CDLL(None).free(dense)
    L.free_flappie_imatrix(im)
    # bundle loader: rejects junk on any box; on a box without a GPU a good bundle is refused with a message, not computed
    L.ffb_model_load.restype = ctypes.c_void_p
    L.ffb_model_load.argtypes = [ctypes.c_char_p, ctypes.c_int]
    bad = tmp_path / "bad.ffbw"; bad.write_bytes(b"FFBW1\0\0\0" + b"\0" * 8)
    assert not L.ffb_model_load(str(bad).encode(), 0) and b"not a weight bundle" in L.ffb_last_error()
    assert not L.ffb_model_load(b"/nonexistent.ffbw", 0)
    if lib.device_count() == 0:
        fm = FlipflopModel.synthetic(KIND_GRU, 64, 4, seed=1)
        good = tmp_path / "r941_native.ffbw"
        fm.save_bundle(str(good))
        assert not L.ffb_model_load(str(good).encode(), 0) and b"not available" in L.ffb_last_error()
        os.environ["FLAPPIE_B200_MODELS"] = str(tmp_path)
        try:
            sig = np.zeros(500, np.float32)
            rt = api.RawTable(None, 500, 0, 500, sig.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
            assert not L.calculate_transitions(rt, 1.0, 0)            # lazy load attempted, refused: NULL, no fallback
        finally:
            del os.environ["FLAPPIE_B200_MODELS"]
