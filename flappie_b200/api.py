"""Python host side above the C ABI of libflappie_b200.so (include/flappie_b200.h).

Mirrors the reference's interface for the hot path: `Library.calculate_transitions`,
`decode_crf_flipflop`, `transpost_crf_flipflop`, `trace_from_posterior` keep the
reference's names and argument meaning (reference src/networks.h:36, src/decode.h:25-38)
and operate on `_Mat` images; `Context.basecall()` is the batched extension used by the
rewired read loop (reference src/flappie.c:364-385).

No arithmetic happens here and there is no fallback: if the shared library is missing
`Library()` raises, and if no CUDA device is present every compute call raises
`FlappieB200Error`.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_bool, c_char_p, c_float, c_int, c_int32, c_int64, c_long, c_size_t, c_uint8, c_uint32, c_void_p
from typing import List, Optional, Sequence

import numpy as np

from .model import FlipflopModel, Mat

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libflappie_b200.so")

FLAG_VITERBI_ONLY = 1
FLAG_WANT_TRACE = 2
FLAG_WANT_TRANS = 4
FLAG_KEEP_LAYERS = 8
FLAG_FP32_SIMT = 16
FLAG_FP32_CONV = 32
FLAG_REVERSE = 64

# enum model_type, reference src/networks.h:18-26
MODEL_ENUM = {"r941_native": 0, "r941_rna002": 1, "r941_5mC": 2, "r103_native": 3, "r10C_pcr": 0, "rle_r941_native": 5}

# every symbol include/flappie_b200.h declares
EXPORTS = [
    "make_flappie_matrix", "free_flappie_matrix", "make_flappie_imatrix", "free_flappie_imatrix",
    "get_flappie_model_type", "flappie_model_string", "flappie_model_description",
    "calculate_transitions", "transpost_crf_flipflop", "decode_crf_flipflop", "trace_from_posterior",
    "exp_activation_inplace", "nbase_from_flipflop_nparam",
    "ffb_device_count", "ffb_last_error", "ffb_version", "ffb_model_create", "ffb_model_destroy",
    "ffb_model_size", "ffb_model_nparam", "ffb_model_stride", "ffb_model_nblock", "ffb_register_model",
    "ffb_create", "ffb_destroy", "ffb_basecall_batch", "ffb_upload", "ffb_forward", "ffb_download", "ffb_sync",
    "ffb_total_blocks", "ffb_launch_count", "ffb_forward_timed", "ffb_debug_fetch", "ffb_debug_group_times", "ffb_emit_bases",
    "ffb_upload_raw", "ffb_basecall_raw_batch", "ffb_submit_batch", "ffb_submit_raw_batch", "ffb_collect",
    "ffb_alloc_pinned", "ffb_free_pinned",
    "decode_crf_runlength", "transpost_crf_runlength", "ffb_emit_runs", "ffb_plan_schedule", "ffb_phred_table",
    "change_positions", "nbase_from_crf_runlength_nparam", "array_from_flappie_imatrix", "ffb_model_load",
    "ffb_submit_raw_begin", "ffb_submit_raw_finish", "ffb_reserve",
]


class FlappieB200Error(RuntimeError):
    pass


class IMat(ctypes.Structure):
    _fields_ = [("nr", c_size_t), ("nrq", c_size_t), ("nc", c_size_t), ("stride", c_size_t),
                ("data", POINTER(c_int32))]


class RawTable(ctypes.Structure):
    """reference src/flappie_structures.h:16-22"""
    _fields_ = [("uuid", c_char_p), ("n", c_size_t), ("start", c_size_t), ("end", c_size_t),
                ("raw", POINTER(c_float))]


class Batch(ctypes.Structure):
    _fields_ = [
        ("signal", POINTER(c_float)), ("sig_off", POINTER(c_int64)), ("n_reads", c_int64),
        ("temperature", c_float), ("flags", c_uint32),
        ("blk_off", POINTER(c_int64)), ("path", POINTER(c_int32)), ("qpath", POINTER(c_float)),
        ("score", POINTER(c_float)), ("trans", POINTER(c_float)), ("tpost", POINTER(c_float)),
        ("trace", POINTER(c_uint8)), ("rle_params", POINTER(c_float)),
        ("bases", POINTER(ctypes.c_char)), ("quals", POINTER(ctypes.c_char)), ("nbases", POINTER(c_int32)),
    ]


class RawBatch(ctypes.Structure):
    """ffb_raw_batch: raw reads + the reference CLI's trimming / normalisation options."""
    _fields_ = [
        ("raw", POINTER(c_float)), ("raw_off", POINTER(c_int64)), ("n_reads", c_int64),
        ("trim_start", c_int64), ("trim_end", c_int64), ("varseg_chunk", c_int64),
        ("varseg_thresh", c_float), ("delta", c_float),
        ("start", POINTER(c_int64)), ("end", POINTER(c_int64)),
    ]


_lib_singleton = None


class Library:
    """The loaded C ABI.  Raises if the shared library has not been built."""

    def __init__(self, path: str = LIB_PATH):
        if not os.path.exists(path):
            raise FlappieB200Error(
                f"{path} is missing: build it with `python -m flappie_b200.build` "
                "(there is no CPU fallback for the hot path)")
        L = self.lib = ctypes.CDLL(path)
        self.path = path
        PM, PI = POINTER(Mat), POINTER(IMat)
        L.make_flappie_matrix.restype = PM; L.make_flappie_matrix.argtypes = [c_size_t, c_size_t]
        L.free_flappie_matrix.restype = PM; L.free_flappie_matrix.argtypes = [PM]
        L.make_flappie_imatrix.restype = PI; L.make_flappie_imatrix.argtypes = [c_size_t, c_size_t]
        L.free_flappie_imatrix.restype = PI; L.free_flappie_imatrix.argtypes = [PI]
        L.get_flappie_model_type.restype = c_int; L.get_flappie_model_type.argtypes = [c_char_p]
        L.flappie_model_string.restype = c_char_p; L.flappie_model_string.argtypes = [c_int]
        L.flappie_model_description.restype = c_char_p; L.flappie_model_description.argtypes = [c_int]
        L.calculate_transitions.restype = PM; L.calculate_transitions.argtypes = [RawTable, c_float, c_int]
        L.transpost_crf_flipflop.restype = PM; L.transpost_crf_flipflop.argtypes = [PM, c_bool]
        L.decode_crf_flipflop.restype = c_float
        L.decode_crf_flipflop.argtypes = [PM, c_bool, POINTER(c_int), POINTER(c_float)]
        L.trace_from_posterior.restype = PI; L.trace_from_posterior.argtypes = [PM]
        L.exp_activation_inplace.restype = None; L.exp_activation_inplace.argtypes = [PM]
        L.nbase_from_flipflop_nparam.restype = c_size_t; L.nbase_from_flipflop_nparam.argtypes = [c_size_t]
        L.decode_crf_runlength.restype = c_float; L.decode_crf_runlength.argtypes = [PM, POINTER(c_int)]
        L.transpost_crf_runlength.restype = PM; L.transpost_crf_runlength.argtypes = [PM]
        L.ffb_emit_runs.restype = c_int64
        L.ffb_emit_runs.argtypes = [POINTER(c_int32), POINTER(c_float), c_int64, c_int, c_char_p, POINTER(c_float),
                                    POINTER(c_float), POINTER(c_int32)]
        L.ffb_device_count.restype = c_int
        L.ffb_last_error.restype = c_char_p
        L.ffb_version.restype = c_char_p
        L.ffb_model_create.restype = c_void_p
        L.ffb_model_create.argtypes = [c_int, c_int, POINTER(POINTER(Mat)), c_int, POINTER(c_int), c_int]
        L.ffb_model_destroy.restype = None; L.ffb_model_destroy.argtypes = [c_void_p]
        for n in ("ffb_model_size", "ffb_model_nparam", "ffb_model_stride"):
            getattr(L, n).restype = c_int; getattr(L, n).argtypes = [c_void_p]
        L.ffb_model_nblock.restype = c_long; L.ffb_model_nblock.argtypes = [c_void_p, c_long]
        L.ffb_register_model.restype = c_int; L.ffb_register_model.argtypes = [c_int, c_void_p]
        L.ffb_create.restype = c_void_p; L.ffb_create.argtypes = [c_void_p, c_void_p]
        L.ffb_destroy.restype = None; L.ffb_destroy.argtypes = [c_void_p]
        for n in ("ffb_basecall_batch", "ffb_upload", "ffb_download"):
            getattr(L, n).restype = c_int; getattr(L, n).argtypes = [c_void_p, POINTER(Batch)]
        for n in ("ffb_submit_batch", "ffb_collect"):
            getattr(L, n).restype = c_int; getattr(L, n).argtypes = [c_void_p, POINTER(Batch)]
        for n in ("ffb_upload_raw", "ffb_basecall_raw_batch", "ffb_submit_raw_batch"):
            getattr(L, n).restype = c_int; getattr(L, n).argtypes = [c_void_p, POINTER(RawBatch), POINTER(Batch)]
        L.ffb_forward.restype = c_int; L.ffb_forward.argtypes = [c_void_p]
        L.ffb_sync.restype = c_int; L.ffb_sync.argtypes = [c_void_p]
        L.ffb_total_blocks.restype = c_int64; L.ffb_total_blocks.argtypes = [c_void_p]
        L.ffb_launch_count.restype = c_int64; L.ffb_launch_count.argtypes = [c_void_p]
        L.ffb_forward_timed.restype = c_int; L.ffb_forward_timed.argtypes = [c_void_p, POINTER(c_float)]
        L.ffb_debug_fetch.restype = c_int64; L.ffb_debug_fetch.argtypes = [c_void_p, c_int, c_void_p, c_int64]
        L.ffb_phred_table.restype = c_int
        L.ffb_phred_table.argtypes = [POINTER(c_float), c_int]
        L.ffb_emit_bases.restype = c_int
        L.ffb_emit_bases.argtypes = [POINTER(c_int32), POINTER(c_float), c_int64, c_int, c_bool, c_char_p, c_char_p]

    @staticmethod
    def get() -> "Library":
        global _lib_singleton
        if _lib_singleton is None:
            _lib_singleton = Library()
        return _lib_singleton

    def last_error(self) -> str:
        return (self.lib.ffb_last_error() or b"").decode()

    def device_count(self) -> int:
        return int(self.lib.ffb_device_count())

    def require_gpu(self) -> None:
        if self.device_count() < 1:
            raise FlappieB200Error("no CUDA device visible: flappie_b200 has no CPU fallback")

    # ---- `_Mat` helpers -----------------------------------------------------------
    def mat_from_rows(self, rows: np.ndarray):
        """[nc][nr] numpy (one row per reference column) -> malloc'ed `_Mat*` (caller frees)."""
        rows = np.ascontiguousarray(rows, np.float32)
        nc, nr = rows.shape
        m = self.lib.make_flappie_matrix(nr, nc)
        if not m:
            raise MemoryError("make_flappie_matrix")
        stride = m.contents.stride
        dst = np.ctypeslib.as_array(m.contents.data, shape=(nc, stride))
        dst[:, :nr] = rows
        return m

    def rows_from_mat(self, m, free: bool = True) -> np.ndarray:
        nr, nc, stride = m.contents.nr, m.contents.nc, m.contents.stride
        out = np.ctypeslib.as_array(m.contents.data, shape=(nc, stride))[:, :nr].copy()
        if free:
            self.lib.free_flappie_matrix(m)
        return out

    # ---- reference-named drop-ins (dense numpy in / out) ---------------------------
    def decode_crf_flipflop(self, trans: np.ndarray, combine_stays: bool = False):
        """trans [T][nr] -> (score, path[T+1], qpath[T+1]); reference src/decode.c:119-204."""
        self.require_gpu()
        T = trans.shape[0]
        tm = self.mat_from_rows(trans)
        path = np.zeros(T + 2, np.int32)
        qpath = np.zeros(T + 2, np.float32)
        score = self.lib.decode_crf_flipflop(tm, combine_stays, path.ctypes.data_as(POINTER(c_int)),
                                             qpath.ctypes.data_as(POINTER(c_float)))
        self.lib.free_flappie_matrix(tm)
        if np.isnan(score):
            raise FlappieB200Error("decode_crf_flipflop failed: " + self.last_error())
        return float(score), path[:T + 1], qpath[:T + 1]

    def transpost_crf_flipflop(self, trans: np.ndarray, return_log: bool = True) -> np.ndarray:
        self.require_gpu()
        tm = self.mat_from_rows(trans)
        out = self.lib.transpost_crf_flipflop(tm, return_log)
        self.lib.free_flappie_matrix(tm)
        if not out:
            raise FlappieB200Error("transpost_crf_flipflop failed: " + self.last_error())
        return self.rows_from_mat(out)

    def trace_from_posterior(self, tpost_prob: np.ndarray) -> np.ndarray:
        self.require_gpu()
        tm = self.mat_from_rows(tpost_prob)
        tr = self.lib.trace_from_posterior(tm)
        self.lib.free_flappie_matrix(tm)
        if not tr:
            raise FlappieB200Error("trace_from_posterior failed: " + self.last_error())
        nr, nc, stride = tr.contents.nr, tr.contents.nc, tr.contents.stride
        out = np.ctypeslib.as_array(tr.contents.data, shape=(nc, stride))[:, :nr].copy()
        self.lib.free_flappie_imatrix(tr)
        return out

    def exp_activation_inplace(self, x: np.ndarray) -> np.ndarray:
        self.require_gpu()
        tm = self.mat_from_rows(x)
        self.lib.exp_activation_inplace(tm)
        return self.rows_from_mat(tm)

    def calculate_transitions(self, signal: np.ndarray, temperature: float, model_enum: int) -> np.ndarray:
        """Per-read drop-in (reference src/networks.c:108-111); needs `Model.register()`."""
        self.require_gpu()
        signal = np.ascontiguousarray(signal, np.float32)
        rt = RawTable(None, signal.shape[0], 0, signal.shape[0], signal.ctypes.data_as(POINTER(c_float)))
        out = self.lib.calculate_transitions(rt, temperature, model_enum)
        if not out:
            return None
        return self.rows_from_mat(out)

    # ---- run-length ("runnie") drop-ins --------------------------------------------
    def decode_crf_runlength(self, param: np.ndarray):
        """param [T][40] -> (score, path[T]); reference src/decode.c:901-984."""
        self.require_gpu()
        T = param.shape[0]
        pm = self.mat_from_rows(param)
        path = np.zeros(T + 2, np.int32)
        score = self.lib.decode_crf_runlength(pm, path.ctypes.data_as(POINTER(c_int)))
        self.lib.free_flappie_matrix(pm)
        if np.isnan(score):
            raise FlappieB200Error("decode_crf_runlength failed: " + self.last_error())
        return float(score), path[:T]

    def transpost_crf_runlength(self, param: np.ndarray) -> np.ndarray:
        self.require_gpu()
        pm = self.mat_from_rows(param)
        out = self.lib.transpost_crf_runlength(pm)
        self.lib.free_flappie_matrix(pm)
        if not out:
            raise FlappieB200Error("transpost_crf_runlength failed: " + self.last_error())
        return self.rows_from_mat(out)

    def emit_runs(self, path: np.ndarray, rle_params: np.ndarray, nbase: int = 4):
        """runnie's run loop (src/runnie.c:279-310): (bases, shape[], scale[], dwell[])"""
        path = np.ascontiguousarray(path, np.int32)
        rle_params = np.ascontiguousarray(rle_params, np.float32)
        T = path.shape[0]
        bases = ctypes.create_string_buffer(T + 2)
        shape = np.zeros(T + 1, np.float32); scale = np.zeros(T + 1, np.float32); dwell = np.zeros(T + 1, np.int32)
        n = self.lib.ffb_emit_runs(path.ctypes.data_as(POINTER(c_int32)), rle_params.ctypes.data_as(POINTER(c_float)), T, nbase,
                                   bases, shape.ctypes.data_as(POINTER(c_float)), scale.ctypes.data_as(POINTER(c_float)),
                                   dwell.ctypes.data_as(POINTER(c_int32)))
        if n < 0:
            raise FlappieB200Error("ffb_emit_runs: bad arguments")
        return bases.raw[:n].decode(), shape[:n], scale[:n], dwell[:n]

    def phred_table(self) -> np.ndarray:
        """qpath thresholds of the quality characters (ascending): char = 33 + #{thresholds <= qpath}"""
        out = np.zeros(128, np.float32)
        n = self.lib.ffb_phred_table(out.ctypes.data_as(POINTER(c_float)), 128)
        return out[:n]

    def emit_bases(self, path: np.ndarray, qpath: np.ndarray, nbase: int, reverse: bool = False):
        path = np.ascontiguousarray(path, np.int32)
        qpath = np.ascontiguousarray(qpath, np.float32)
        nblock = path.shape[0] - 1
        bc = ctypes.create_string_buffer(nblock + 2)
        ql = ctypes.create_string_buffer(nblock + 2)
        n = self.lib.ffb_emit_bases(path.ctypes.data_as(POINTER(c_int32)), qpath.ctypes.data_as(POINTER(c_float)),
                                    nblock, nbase, reverse, bc, ql)
        if n < 0:
            raise FlappieB200Error("ffb_emit_bases: bad arguments")
        return bc.raw[:n].decode(), ql.raw[:n].decode()


class Model:
    """Device-resident weight arena built from the reference's `_Mat` bundle."""

    def __init__(self, fm: FlipflopModel, device: int = 0, lib: Optional[Library] = None):
        self.lib = lib or Library.get()
        self.lib.require_gpu()
        self.fm = fm
        mats, keep = fm.to_mat_bundle()
        arr = (POINTER(Mat) * len(mats))(*[ctypes.pointer(m) for m in mats])
        strides = (c_int * len(fm.conv_stride))(*fm.conv_stride)
        kind = 2 if getattr(fm, "head", "flipflop") == "runlength" else fm.kind      # FFB_KIND_RUNLENGTH
        self.handle = self.lib.lib.ffb_model_create(device, kind, arr, len(mats), strides, len(fm.conv_stride))
        del keep
        if not self.handle:
            raise FlappieB200Error("ffb_model_create failed: " + self.lib.last_error())
        self.device = device

    def register(self, model_name: str) -> None:
        r = self.lib.lib.ffb_register_model(MODEL_ENUM[model_name], self.handle)
        if r != 0:
            raise FlappieB200Error("ffb_register_model failed")

    def nblock(self, nsample: int) -> int:
        return int(self.lib.lib.ffb_model_nblock(self.handle, nsample))

    def close(self):
        if self.handle:
            self.lib.lib.ffb_model_destroy(self.handle)
            self.handle = None


class BatchResult:
    def __init__(self, n_reads, blk_off, path, qpath, score, trans, tpost, trace, nstate, nparam):
        self.n_reads, self.blk_off, self.path, self.qpath, self.score = n_reads, blk_off, path, qpath, score
        self.trans, self.tpost, self.trace, self.nstate, self.nparam = trans, tpost, trace, nstate, nparam
        self.rle_params = None
        self.bases = self.quals = self.nbases = None

    def read_bases(self, i: int):
        """(basecall, quality) strings of read i as emitted on the device"""
        s = int(self.blk_off[i]) + i
        k = int(self.nbases[i])
        return self.bases[s:s + k].tobytes().decode("ascii"), self.quals[s:s + k].tobytes().decode("ascii")

    def nblock(self, i: int) -> int:
        return int(self.blk_off[i + 1] - self.blk_off[i])

    def read_path(self, i: int):
        s = int(self.blk_off[i]) + i
        return self.path[s:s + self.nblock(i) + 1], self.qpath[s:s + self.nblock(i) + 1]

    def read_trans(self, i: int):
        return self.trans[int(self.blk_off[i]):int(self.blk_off[i + 1])]

    def read_tpost(self, i: int):
        return self.tpost[int(self.blk_off[i]):int(self.blk_off[i + 1])]

    def read_rle(self, i: int):
        """run-length models: (states[T], shape/scale rows [T][8]) of read i"""
        s = int(self.blk_off[i]) + i
        return self.path[s:s + self.nblock(i)], self.rle_params[int(self.blk_off[i]):int(self.blk_off[i + 1])]

    def read_trace(self, i: int):
        s = int(self.blk_off[i]) + i
        return self.trace[s:s + self.nblock(i) + 1]


class Context:
    """One per (model, stream): grow-only HBM workspaces + the batch plan."""

    def __init__(self, model: Model, stream: int = 0):
        self.model = model
        self.lib = model.lib
        self.handle = self.lib.lib.ffb_create(model.handle, c_void_p(stream) if stream else None)
        if not self.handle:
            raise FlappieB200Error("ffb_create failed: " + self.lib.last_error())
        self._keep = None

    def close(self):
        if self.handle:
            self.lib.lib.ffb_destroy(self.handle)
            self.handle = None

    def _check(self, r: int, what: str):
        if r != 0:
            raise FlappieB200Error(f"{what} failed ({r}): {self.lib.last_error()}")

    def make_batch(self, signal: np.ndarray, sig_off: np.ndarray, temperature=1.0, flags=0,
                   out: Optional[dict] = None, emit: bool = False, want_path: bool = True):
        """Build the C `ffb_batch` over caller-owned numpy (or pinned torch->numpy) buffers."""
        fm = self.model.fm
        n = sig_off.shape[0] - 1
        assert signal.dtype == np.float32 and sig_off.dtype == np.int64
        tot_blocks = 0
        for i in range(n):
            t = fm.nblock(int(sig_off[i + 1] - sig_off[i]))
            tot_blocks += max(t, 0)
        o = out if out is not None else {}
        o.setdefault("blk_off", np.zeros(n + 1, np.int64))
        if want_path:       # with device-side emission the path / qpath need not come back at all
            o.setdefault("path", np.zeros(tot_blocks + n, np.int32))
            o.setdefault("qpath", np.zeros(tot_blocks + n, np.float32))
        o.setdefault("score", np.zeros(max(n, 1), np.float32))
        if flags & FLAG_WANT_TRANS:
            o.setdefault("trans", np.zeros((tot_blocks, fm.nparam), np.float32))
            if not flags & FLAG_VITERBI_ONLY:
                o.setdefault("tpost", np.zeros((tot_blocks, fm.nparam), np.float32))
        if flags & FLAG_WANT_TRACE:
            o.setdefault("trace", np.zeros((tot_blocks + n, fm.nstate), np.uint8))
        if getattr(fm, "head", "flipflop") == "runlength":
            o.setdefault("rle_params", np.zeros((tot_blocks, 8), np.float32))
        if emit:
            # device-side emission (emit.cu): chars per read at blk_off[n] + n, like path
            o.setdefault("bases", np.zeros(tot_blocks + n, np.uint8))
            o.setdefault("quals", np.zeros(tot_blocks + n, np.uint8))
            o.setdefault("nbases", np.zeros(max(n, 1), np.int32))

        def p(name, ct):
            a = o.get(name)
            return a.ctypes.data_as(POINTER(ct)) if a is not None else None

        b = Batch(signal.ctypes.data_as(POINTER(c_float)), sig_off.ctypes.data_as(POINTER(c_int64)), n,
                  temperature, flags, p("blk_off", c_int64), p("path", c_int32), p("qpath", c_float),
                  p("score", c_float), p("trans", c_float), p("tpost", c_float), p("trace", c_uint8), p("rle_params", c_float),
                  p("bases", ctypes.c_char), p("quals", ctypes.c_char), p("nbases", c_int32))
        self._keep = (signal, sig_off, o)
        return b, o

    def basecall(self, reads: Sequence[np.ndarray], temperature: float = 1.0, viterbi_only: bool = False,
                 want_trace: bool = False, want_trans: bool = False, keep_layers: bool = False,
                 fp32_simt: bool = False, emit: bool = False, reverse: bool = False) -> BatchResult:
        """Whole hot path for a list of already-normalised reads (host numpy arrays)."""
        fm = self.model.fm
        n = len(reads)
        lens = np.array([len(r) for r in reads], np.int64)
        sig_off = np.zeros(n + 1, np.int64)
        np.cumsum(lens, out=sig_off[1:])
        signal = np.concatenate([np.asarray(r, np.float32) for r in reads]) if n else np.zeros(1, np.float32)
        flags = (FLAG_VITERBI_ONLY if viterbi_only else 0) | (FLAG_WANT_TRACE if want_trace else 0) | \
                (FLAG_WANT_TRANS if want_trans else 0) | (FLAG_KEEP_LAYERS if keep_layers else 0) | \
                (FLAG_FP32_SIMT if fp32_simt else 0) | (FLAG_REVERSE if reverse else 0)
        b, o = self.make_batch(signal, sig_off, temperature, flags, emit=emit)
        self._check(self.lib.lib.ffb_basecall_batch(self.handle, ctypes.byref(b)), "ffb_basecall_batch")
        res = BatchResult(n, o["blk_off"], o["path"], o["qpath"], o["score"], o.get("trans"), o.get("tpost"),
                          o.get("trace"), fm.nstate, fm.nparam)
        res.rle_params = o.get("rle_params")
        res.bases, res.quals, res.nbases = o.get("bases"), o.get("quals"), o.get("nbases")
        return res

    def make_raw_batch(self, raw: np.ndarray, raw_off: np.ndarray, trim=(200, 10), segmentation=(100, 0.0),
                       delta: float = 0.0):
        """C `ffb_raw_batch` over caller-owned buffers; defaults = the reference CLI's (src/flappie.c:100-110)."""
        n = raw_off.shape[0] - 1
        assert raw.dtype == np.float32 and raw_off.dtype == np.int64
        start = np.zeros(max(n, 1), np.int64)
        end = np.zeros(max(n, 1), np.int64)
        rb = RawBatch(raw.ctypes.data_as(POINTER(c_float)), raw_off.ctypes.data_as(POINTER(c_int64)), n,
                      int(trim[0]), int(trim[1]), int(segmentation[0]), float(segmentation[1]), float(delta),
                      start.ctypes.data_as(POINTER(c_int64)), end.ctypes.data_as(POINTER(c_int64)))
        self._keep_raw = (raw, raw_off, start, end)
        return rb, start, end

    def basecall_raw(self, raws: Sequence[np.ndarray], temperature: float = 1.0, viterbi_only: bool = False,
                     want_trace: bool = False, want_trans: bool = False, trim=(200, 10), segmentation=(100, 0.0),
                     delta: float = 0.0, emit: bool = False, reverse: bool = False) -> BatchResult:
        """calculate_post from the raw signal on (reference src/flappie.c:245-316) for a list of raw reads
        (pA floats): trimming + normalisation on the device, then the whole hot path.  The result carries
        `start` / `end` (kept range per read)."""
        fm = self.model.fm
        n = len(raws)
        lens = np.array([len(r) for r in raws], np.int64)
        raw_off = np.zeros(n + 1, np.int64)
        np.cumsum(lens, out=raw_off[1:])
        raw = np.concatenate([np.asarray(r, np.float32) for r in raws]) if n else np.zeros(1, np.float32)
        flags = (FLAG_VITERBI_ONLY if viterbi_only else 0) | (FLAG_WANT_TRACE if want_trace else 0) | \
                (FLAG_WANT_TRANS if want_trans else 0) | (FLAG_REVERSE if reverse else 0)
        # outputs sized for the untrimmed lengths (upper bound on the block count)
        b, o = self.make_batch(raw, raw_off, temperature, flags, emit=emit)
        rb, start, end = self.make_raw_batch(raw, raw_off, trim, segmentation, delta)
        self._check(self.lib.lib.ffb_basecall_raw_batch(self.handle, ctypes.byref(rb), ctypes.byref(b)), "ffb_basecall_raw_batch")
        res = BatchResult(n, o["blk_off"], o["path"], o["qpath"], o["score"], o.get("trans"), o.get("tpost"),
                          o.get("trace"), fm.nstate, fm.nparam)
        res.start, res.end = start[:n], end[:n]
        res.rle_params = o.get("rle_params")
        res.bases, res.quals, res.nbases = o.get("bases"), o.get("quals"), o.get("nbases")
        return res

    def fetch_signal(self, n_samples: int) -> np.ndarray:
        """The normalised signal the network read (concatenated kept ranges of the last batch)."""
        out = np.zeros(max(n_samples, 1), np.float32)
        r = self.lib.lib.ffb_debug_fetch(self.handle, 8, out.ctypes.data_as(c_void_p), 4 * n_samples)
        if r < 0:
            raise FlappieB200Error("ffb_debug_fetch failed: " + self.lib.last_error())
        return out[:n_samples]

    def submit_raw(self, rb: RawBatch, b: Batch):
        """enqueue upload + kernels + D2H for one raw batch; results are valid after collect(b)"""
        self._check(self.lib.lib.ffb_submit_raw_batch(self.handle, ctypes.byref(rb), ctypes.byref(b)), "ffb_submit_raw_batch")

    def submit_raw_begin(self, rb: RawBatch, b: Batch):
        """first half of submit_raw: raw H2D + trimming / normalisation kernels enqueued, returns at once"""
        self._check(self.lib.lib.ffb_submit_raw_begin(self.handle, ctypes.byref(rb), ctypes.byref(b)), "ffb_submit_raw_begin")

    def submit_raw_finish(self, b: Batch):
        """second half: waits for the kept ranges, plans the batch and enqueues the network, decode and D2H"""
        self._check(self.lib.lib.ffb_submit_raw_finish(self.handle, ctypes.byref(b)), "ffb_submit_raw_finish")

    def reserve(self, n_reads: int, samples_per_read: int, flags: int = 0):
        """size every device workspace and the pinned plan arena for batches up to this shape before the first one"""
        self._check(self.lib.lib.ffb_reserve(self.handle, c_int64(n_reads), c_int64(samples_per_read), c_uint32(flags)), "ffb_reserve")

    def submit(self, b: Batch):
        self._check(self.lib.lib.ffb_submit_batch(self.handle, ctypes.byref(b)), "ffb_submit_batch")

    def collect(self, b: Batch):
        self._check(self.lib.lib.ffb_collect(self.handle, ctypes.byref(b)), "ffb_collect")

    def upload(self, b: Batch):
        self._check(self.lib.lib.ffb_upload(self.handle, ctypes.byref(b)), "ffb_upload")

    def forward(self):
        self._check(self.lib.lib.ffb_forward(self.handle), "ffb_forward")

    def download(self, b: Batch):
        self._check(self.lib.lib.ffb_download(self.handle, ctypes.byref(b)), "ffb_download")

    def sync(self):
        self._check(self.lib.lib.ffb_sync(self.handle), "ffb_sync")

    def forward_timed(self):
        ms = (c_float * 8)()
        self._check(self.lib.lib.ffb_forward_timed(self.handle, ms), "ffb_forward_timed")
        return dict(conv=ms[0], gemm=ms[1], rnn=ms[2], out=ms[3], decode=ms[4], total=ms[5], recurrent_layers=ms[6])

    def launch_count(self) -> int:
        return int(self.lib.lib.ffb_launch_count(self.handle))

    def total_blocks(self) -> int:
        return int(self.lib.lib.ffb_total_blocks(self.handle))

    def fetch_layer(self, what: int) -> np.ndarray:
        """what: 0 conv output, 1..5 recurrent layers (both need keep_layers), 6 trans."""
        fm = self.model.fm
        Tt = self.total_blocks()
        width = fm.nparam if what == 6 else fm.size
        out = np.zeros((Tt, width), np.float32)
        r = self.lib.lib.ffb_debug_fetch(self.handle, what, out.ctypes.data_as(c_void_p), out.nbytes)
        if r < 0:
            raise FlappieB200Error("ffb_debug_fetch failed: " + self.lib.last_error())
        return out

    def fetch_logz(self, n_reads: int) -> np.ndarray:
        out = np.zeros(n_reads, np.float64)
        r = self.lib.lib.ffb_debug_fetch(self.handle, 7, out.ctypes.data_as(c_void_p), out.nbytes)
        if r < 0:
            raise FlappieB200Error("ffb_debug_fetch failed: " + self.lib.last_error())
        return out
