"""ctypes access to the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

* `Oracle`  -- oracle/_build/liboracle.so, the plain-C restatement (flappie_oracle.c)
* `Ref`     -- oracle/_ref/libflappie_ref.so, the reference's own sources compiled
               unmodified from /root/reference/src plus oracle/ref_driver.c

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; nothing under flappie_b200/ does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_bool, c_char_p, c_double, c_float, c_int, c_int32, c_long, c_size_t, c_void_p

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libflappie_ref.so")

ACT_TANH, ACT_SWISH, ACT_NONE = 0, 1, 2

_f32p = POINTER(c_float)
_i32p = POINTER(c_int32)


def _fp(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_f32p)


def build(ref: bool = True) -> None:
    """Compile the checkers (oracle always; _ref only when /root/reference exists)."""
    subprocess.run(["make", "-C", HERE, "oracle"], check=True, capture_output=True)
    if ref and os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)
        # the link proof: the reference's own main() against libflappie_b200.so (needs the product library to exist)
        if os.path.exists(os.path.join(HERE, "..", "flappie_b200", "csrc", "libflappie_b200.so")):
            subprocess.run(["make", "-C", HERE, "linkproof"], check=True, capture_output=True)


class ConvTerm(ctypes.Structure):
    _fields_ = [("col", c_int32), ("x_start", c_int32), ("tap_lo", c_int32), ("ntap", c_int32)]


class _OModel(ctypes.Structure):
    _fields_ = [
        ("kind", c_int), ("nconv", c_int),
        ("conv_nf", c_int * 3), ("conv_nfilter", c_int * 3), ("conv_winlen", c_int * 3),
        ("conv_stride", c_int * 3),
        ("conv_W", _f32p * 3), ("conv_b", _f32p * 3),
        ("size", c_int),
        ("iW", _f32p * 5), ("sW", _f32p * 5), ("b", _f32p * 5),
        ("nparam", c_int), ("FF_W", _f32p), ("FF_b", _f32p),
    ]


class Oracle:
    """The plain-C restatement."""

    def __init__(self, path: str = ORACLE_SO):
        if not os.path.exists(path):
            build(ref=False)
        L = self.lib = ctypes.CDLL(path)
        L.ffo_conv_plan.restype = c_long
        L.ffo_conv_plan.argtypes = [c_int, c_int, c_int, POINTER(ConvTerm), c_long, POINTER(c_int)]
        L.ffo_convolution.restype = c_int
        L.ffo_convolution.argtypes = [_f32p, c_int, c_int, _f32p, _f32p, c_int, c_int, c_int, c_int, _f32p]
        L.ffo_affine.argtypes = [_f32p, c_int, c_int, _f32p, _f32p, c_int, _f32p]
        L.ffo_grumod.argtypes = [_f32p, c_int, c_int, _f32p, c_int, _f32p]
        L.ffo_lstm.argtypes = [_f32p, c_int, c_int, _f32p, c_int, _f32p]
        L.ffo_globalnorm_flipflop.argtypes = [_f32p, c_int, c_int, _f32p, _f32p, c_int, c_float, _f32p, POINTER(c_double)]
        L.ffo_decode_crf_flipflop.restype = c_float
        L.ffo_decode_crf_flipflop.argtypes = [_f32p, c_int, c_int, POINTER(c_int), _f32p]
        L.ffo_transpost_crf_flipflop.restype = c_int
        L.ffo_transpost_crf_flipflop.argtypes = [_f32p, c_int, c_int, c_int, _f32p]
        L.ffo_trace_from_posterior.argtypes = [_f32p, c_int, c_int, _i32p]
        L.ffo_emit_bases.restype = c_int
        L.ffo_emit_bases.argtypes = [POINTER(c_int), _f32p, c_int, c_int, c_char_p, c_char_p]
        L.ffo_transitions.restype = c_int
        L.ffo_transitions.argtypes = [POINTER(_OModel), _f32p, c_int, c_float, _f32p, _f32p, POINTER(_f32p)]
        L.ffo_basecall.restype = c_int
        L.ffo_basecall.argtypes = [POINTER(_OModel), _f32p, c_int, c_float, c_int, c_char_p, c_char_p,
                                   POINTER(c_float), POINTER(c_int), _f32p, _i32p]
        L.ffo_nblock.restype = c_int
        L.ffo_nblock.argtypes = [POINTER(_OModel), c_int]

    # -- pieces -----------------------------------------------------------------
    def conv_plan(self, T, winlen, stride):
        ncol = c_int(0)
        n = self.lib.ffo_conv_plan(T, winlen, stride, None, 0, ctypes.byref(ncol))
        if n < 0:
            return None, 0
        terms = (ConvTerm * max(n, 1))()
        self.lib.ffo_conv_plan(T, winlen, stride, terms, n, ctypes.byref(ncol))
        return [(t.col, t.x_start, t.tap_lo, t.ntap) for t in terms[:n]], ncol.value

    def convolution(self, x, W, b, stride, act):
        """x [T][nf], W [nfilter][winlen][nf] -> [To][nfilter] or None."""
        x = np.ascontiguousarray(x, np.float32)
        W = np.ascontiguousarray(W, np.float32)
        b = np.ascontiguousarray(b, np.float32)
        T, nf = x.shape
        nfilter, winlen, nf2 = W.shape
        assert nf == nf2
        To = (T + stride - 1) // stride
        out = np.zeros((To, nfilter), np.float32)
        r = self.lib.ffo_convolution(_fp(x), T, nf, _fp(W), _fp(b), nfilter, winlen, stride, act, _fp(out))
        return out if r >= 0 else None

    def affine(self, X, W, b):
        X = np.ascontiguousarray(X, np.float32); W = np.ascontiguousarray(W, np.float32)
        b = np.ascontiguousarray(b, np.float32)
        out = np.zeros((X.shape[0], W.shape[0]), np.float32)
        self.lib.ffo_affine(_fp(X), X.shape[0], X.shape[1], _fp(W), _fp(b), W.shape[0], _fp(out))
        return out

    def grumod(self, Xin, sW, backward):
        Xin = np.ascontiguousarray(Xin, np.float32); sW = np.ascontiguousarray(sW, np.float32)
        T, S = Xin.shape[0], sW.shape[1]
        out = np.zeros((T, S), np.float32)
        self.lib.ffo_grumod(_fp(Xin), T, S, _fp(sW), int(backward), _fp(out))
        return out

    def lstm(self, Xin, sW, backward):
        Xin = np.ascontiguousarray(Xin, np.float32); sW = np.ascontiguousarray(sW, np.float32)
        T, S = Xin.shape[0], sW.shape[1]
        out = np.zeros((T, S), np.float32)
        self.lib.ffo_lstm(_fp(Xin), T, S, _fp(sW), int(backward), _fp(out))
        return out

    def globalnorm(self, h, W, b, temperature=1.0):
        h = np.ascontiguousarray(h, np.float32); W = np.ascontiguousarray(W, np.float32)
        b = np.ascontiguousarray(b, np.float32)
        out = np.zeros((h.shape[0], W.shape[0]), np.float32)
        logz = c_double(0)
        self.lib.ffo_globalnorm_flipflop(_fp(h), h.shape[0], h.shape[1], _fp(W), _fp(b), W.shape[0],
                                         temperature, _fp(out), ctypes.byref(logz))
        return out, logz.value

    def viterbi(self, trans):
        trans = np.ascontiguousarray(trans, np.float32)
        T, nr = trans.shape
        path = np.zeros(T + 1, np.int32)
        qpath = np.zeros(T + 1, np.float32)
        score = self.lib.ffo_decode_crf_flipflop(_fp(trans), T, nr, path.ctypes.data_as(POINTER(c_int)), _fp(qpath))
        return score, path, qpath

    def transpost(self, trans, return_log=True):
        trans = np.ascontiguousarray(trans, np.float32)
        out = np.zeros_like(trans)
        self.lib.ffo_transpost_crf_flipflop(_fp(trans), trans.shape[0], trans.shape[1], int(return_log), _fp(out))
        return out

    def trace(self, tpost_prob):
        tpost_prob = np.ascontiguousarray(tpost_prob, np.float32)
        T, nr = tpost_prob.shape
        nbase = int(round((-1.0 + np.sqrt(1.0 + 2.0 * nr)) / 2.0))
        out = np.zeros((T + 1, 2 * nbase), np.int32)
        self.lib.ffo_trace_from_posterior(_fp(tpost_prob), T, nr, out.ctypes.data_as(_i32p))
        return out

    def emit_bases(self, path, qpath, nbase):
        path = np.ascontiguousarray(path, np.int32); qpath = np.ascontiguousarray(qpath, np.float32)
        nblock = path.shape[0] - 1
        bc = ctypes.create_string_buffer(nblock + 2)
        ql = ctypes.create_string_buffer(nblock + 2)
        n = self.lib.ffo_emit_bases(path.ctypes.data_as(POINTER(c_int)), _fp(qpath), nblock, nbase, bc, ql)
        return bc.raw[:n].decode(), ql.raw[:n].decode()

    # -- run-length ("runnie") head ----------------------------------------------
    def globalnorm_runlength(self, h, W, b, temperature=1.0):
        h = np.ascontiguousarray(h, np.float32); W = np.ascontiguousarray(W, np.float32)
        b = np.ascontiguousarray(b, np.float32)
        out = np.zeros((h.shape[0], W.shape[0]), np.float32)
        logz = c_double(0)
        self.lib.ffo_globalnorm_runlength.argtypes = [_f32p, c_int, c_int, _f32p, _f32p, c_int, c_float, _f32p, POINTER(c_double)]
        self.lib.ffo_globalnorm_runlength(_fp(h), h.shape[0], h.shape[1], _fp(W), _fp(b), W.shape[0], temperature,
                                          _fp(out), ctypes.byref(logz))
        return out, logz.value

    def rle_viterbi(self, param):
        param = np.ascontiguousarray(param, np.float32)
        T, nr = param.shape
        path = np.zeros(max(T, 1), np.int32)
        self.lib.ffo_decode_crf_runlength.restype = c_float
        self.lib.ffo_decode_crf_runlength.argtypes = [_f32p, c_int, c_int, POINTER(c_int)]
        score = self.lib.ffo_decode_crf_runlength(_fp(param), T, nr, path.ctypes.data_as(POINTER(c_int)))
        return score, path[:T]

    def rle_transpost(self, param):
        param = np.ascontiguousarray(param, np.float32)
        out = np.zeros_like(param)
        self.lib.ffo_transpost_crf_runlength.argtypes = [_f32p, c_int, c_int, _f32p]
        self.lib.ffo_transpost_crf_runlength(_fp(param), param.shape[0], param.shape[1], _fp(out))
        return out

    def emit_runs(self, path, post):
        """runnie's run loop (runnie.c:279-310) -> (bases str, shape[], scale[], dwell[])"""
        path = np.ascontiguousarray(path, np.int32); post = np.ascontiguousarray(post, np.float32)
        T, nr = post.shape
        bases = ctypes.create_string_buffer(T + 1)
        shape = np.zeros(max(T, 1), np.float32); scale = np.zeros(max(T, 1), np.float32); dwell = np.zeros(max(T, 1), np.int32)
        self.lib.ffo_emit_runs.restype = c_int
        self.lib.ffo_emit_runs.argtypes = [POINTER(c_int), _f32p, c_int, c_int, c_char_p, _f32p, _f32p, POINTER(c_int)]
        n = self.lib.ffo_emit_runs(path.ctypes.data_as(POINTER(c_int)), _fp(post), T, nr, bases, _fp(shape), _fp(scale),
                                   dwell.ctypes.data_as(POINTER(c_int)))
        return bases.raw[:n].decode(), shape[:n], scale[:n], dwell[:n]

    def runlength_transitions(self, m, signal, temperature=1.0):
        """runlength5_guppy_transitions (networks.c:675-722): the LSTM stack of the model, then globalnorm_runlengthV2"""
        r = self.transitions(m, signal, temperature, want_layers=True)
        if r is None:
            return None
        return self.globalnorm_runlength(r[2][4], m.ff_W, m.ff_b, temperature)[0]

    # -- whole model ------------------------------------------------------------
    def _model(self, m):
        om = _OModel()
        keep = []
        om.kind = m.kind
        om.nconv = len(m.conv_W)
        for i, (w, bb) in enumerate(zip(m.conv_W, m.conv_b)):
            w = np.ascontiguousarray(w, np.float32); bb = np.ascontiguousarray(bb, np.float32)
            keep += [w, bb]
            om.conv_nfilter[i], om.conv_winlen[i], om.conv_nf[i] = w.shape
            om.conv_stride[i] = m.conv_stride[i]
            om.conv_W[i] = _fp(w); om.conv_b[i] = _fp(bb)
        om.size = m.size
        for i in range(5):
            arrs = [np.ascontiguousarray(a, np.float32) for a in (m.iW[i], m.sW[i], m.b[i])]
            keep += arrs
            om.iW[i], om.sW[i], om.b[i] = (_fp(a) for a in arrs)
        fw = np.ascontiguousarray(m.ff_W, np.float32); fb = np.ascontiguousarray(m.ff_b, np.float32)
        keep += [fw, fb]
        om.nparam = fw.shape[0]
        om.FF_W, om.FF_b = _fp(fw), _fp(fb)
        return om, keep

    def transitions(self, m, signal, temperature=1.0, want_layers=False):
        om, keep = self._model(m)
        signal = np.ascontiguousarray(signal, np.float32)
        T = self.lib.ffo_nblock(ctypes.byref(om), signal.shape[0])
        if T < 0:
            return None
        trans = np.zeros((T, m.nparam), np.float32)
        conv = np.zeros((T, m.conv_W[-1].shape[0]), np.float32)
        layers = [np.zeros((T, m.size), np.float32) for _ in range(5)]
        lp = (_f32p * 5)(*[_fp(a) for a in layers])
        r = self.lib.ffo_transitions(ctypes.byref(om), _fp(signal), signal.shape[0], temperature, _fp(trans),
                                     _fp(conv), lp)
        if r < 0:
            return None
        return (trans, conv, layers) if want_layers else trans

    def basecall(self, m, signal, temperature=1.0, viterbi_only=False, want_trace=False):
        om, keep = self._model(m)
        signal = np.ascontiguousarray(signal, np.float32)
        T = self.lib.ffo_nblock(ctypes.byref(om), signal.shape[0])
        if T < 0:
            return None
        bc = ctypes.create_string_buffer(T + 2); ql = ctypes.create_string_buffer(T + 2)
        score = c_float(0)
        path = np.zeros(T + 2, np.int32); qpath = np.zeros(T + 2, np.float32)
        trace = np.zeros((T + 1, m.nstate), np.int32) if want_trace else None
        n = self.lib.ffo_basecall(ctypes.byref(om), _fp(signal), signal.shape[0], temperature, int(viterbi_only),
                                  bc, ql, ctypes.byref(score), path.ctypes.data_as(POINTER(c_int)), _fp(qpath),
                                  trace.ctypes.data_as(_i32p) if want_trace else None)
        if n < 0:
            return None
        return dict(basecall=bc.raw[:n].decode(), quality=ql.raw[:n].decode(), score=score.value,
                    path=path[:T + 1], qpath=qpath[:T + 1], trace=trace, nblock=T)


# ---------------------------------------------------------------------------------
class _RMat(ctypes.Structure):
    _fields_ = [("nr", c_size_t), ("nrq", c_size_t), ("nc", c_size_t), ("stride", c_size_t), ("data", _f32p)]


class _RModel(ctypes.Structure):
    _fields_ = [
        ("kind", c_int), ("nconv", c_int),
        ("conv_W", POINTER(_RMat) * 3), ("conv_b", POINTER(_RMat) * 3), ("conv_stride", c_int * 3),
        ("iW", POINTER(_RMat) * 5), ("sW", POINTER(_RMat) * 5), ("b", POINTER(_RMat) * 5),
        ("FF_W", POINTER(_RMat)), ("FF_b", POINTER(_RMat)),
    ]


class Ref:
    """The reference's own compiled code (oracle/_ref)."""

    def __init__(self, path: str = REF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = self.lib = ctypes.CDLL(path)
        P = POINTER(_RMat)
        self.P = P
        L.ffref_mat_from_dense.restype = P
        L.ffref_mat_from_dense.argtypes = [_f32p, c_size_t, c_size_t]
        L.ffref_mat_to_dense.argtypes = [P, _f32p]
        L.free_flappie_matrix.restype = P
        L.free_flappie_matrix.argtypes = [P]
        L.convolution.restype = P
        L.convolution.argtypes = [P, P, P, c_size_t, P]
        L.tanh_activation_inplace.argtypes = [P]
        L.swish_activation_inplace.argtypes = [P]
        L.exp_activation_inplace.argtypes = [P]
        L.feedforward_linear.restype = P
        L.feedforward_linear.argtypes = [P, P, P, P]
        for n in ("grumod_forward", "grumod_backward", "lstm_forward", "lstm_backward"):
            getattr(L, n).restype = P
            getattr(L, n).argtypes = [P, P, P]
        L.globalnorm_flipflop.restype = P
        L.globalnorm_flipflop.argtypes = [P, P, P, c_float, P]
        L.crf_manystay_partition_function.restype = c_double
        L.crf_manystay_partition_function.argtypes = [P]
        L.decode_crf_flipflop.restype = c_float
        L.decode_crf_flipflop.argtypes = [P, c_bool, POINTER(c_int), _f32p]
        L.transpost_crf_flipflop.restype = P
        L.transpost_crf_flipflop.argtypes = [P, c_bool]
        L.trace_from_posterior.restype = c_void_p
        L.trace_from_posterior.argtypes = [P]
        L.ffref_imat_to_dense.argtypes = [c_void_p, _i32p]
        L.free_flappie_imatrix.restype = c_void_p
        L.free_flappie_imatrix.argtypes = [c_void_p]
        L.medmad_normalise_array.argtypes = [_f32p, c_size_t]
        L.ffref_transitions.restype = P
        L.ffref_transitions.argtypes = [POINTER(_RModel), _f32p, c_size_t, c_float, POINTER(P)]
        L.ffref_decode.restype = c_long
        L.ffref_decode.argtypes = [P, c_bool, POINTER(c_int), _f32p, c_char_p, c_char_p, POINTER(c_float), _i32p, _f32p]
        L.ffref_basecall.restype = c_long
        L.ffref_basecall.argtypes = [POINTER(_RModel), _f32p, c_size_t, c_float, c_bool, c_char_p, c_char_p, POINTER(c_float)]
        try:
            L.openblas_set_num_threads.argtypes = [c_int]
            L.openblas_set_num_threads(1)   # reference README.md:66-67
        except AttributeError:
            pass
        bi = os.path.join(os.path.dirname(path), "BUILDINFO")
        self.buildinfo = open(bi).read().strip() if os.path.exists(bi) else "unknown"

    # dense [nc][nr] (row per flappie column) <-> reference _Mat
    def mat(self, cols: np.ndarray):
        cols = np.ascontiguousarray(cols, np.float32)
        nc, nr = cols.shape
        return self.lib.ffref_mat_from_dense(_fp(cols), nr, nc)

    def conv_mat(self, W: np.ndarray):
        """[nfilter][winlen][nf] -> conv filter _Mat with the exporter's padding."""
        nfilter, winlen, nf = W.shape
        nf4 = 4 * ((nf + 3) // 4)
        nr = nf4 * winlen - nf4 + nf
        dense = np.zeros((nfilter, nr), np.float32)
        for k in range(winlen):
            dense[:, k * nf4:k * nf4 + nf] = W[:, k, :]
        return self.mat(dense)

    def dense(self, m, free=True) -> np.ndarray:
        nr, nc = m.contents.nr, m.contents.nc
        out = np.zeros((nc, nr), np.float32)
        self.lib.ffref_mat_to_dense(m, _fp(out))
        if free:
            self.lib.free_flappie_matrix(m)
        return out

    def free(self, *ms):
        for m in ms:
            self.lib.free_flappie_matrix(m)

    # -- pieces, dense in / dense out ------------------------------------------
    def convolution(self, x, W, b, stride, act):
        X = self.mat(x); Wm = self.conv_mat(np.asarray(W, np.float32)); bm = self.mat(np.asarray(b, np.float32).reshape(1, -1))
        C = self.lib.convolution(X, Wm, bm, stride, None)
        if act == ACT_TANH:
            self.lib.tanh_activation_inplace(C)
        elif act == ACT_SWISH:
            self.lib.swish_activation_inplace(C)
        out = self.dense(C)
        self.free(X, Wm, bm)
        return out

    def affine(self, X, W, b):
        Xm = self.mat(X); Wm = self.mat(W); bm = self.mat(np.asarray(b, np.float32).reshape(1, -1))
        out = self.dense(self.lib.feedforward_linear(Xm, Wm, bm, None))
        self.free(Xm, Wm, bm)
        return out

    def _rnn(self, fn, Xin, sW):
        Xm = self.mat(Xin); Wm = self.mat(sW)
        out = self.dense(getattr(self.lib, fn)(Xm, Wm, None))
        self.free(Xm, Wm)
        return out

    def grumod(self, Xin, sW, backward):
        return self._rnn("grumod_backward" if backward else "grumod_forward", Xin, sW)

    def lstm(self, Xin, sW, backward):
        return self._rnn("lstm_backward" if backward else "lstm_forward", Xin, sW)

    def globalnorm(self, h, W, b, temperature=1.0):
        hm = self.mat(h); Wm = self.mat(W); bm = self.mat(np.asarray(b, np.float32).reshape(1, -1))
        out = self.dense(self.lib.globalnorm_flipflop(hm, Wm, bm, temperature, None))
        self.free(hm, Wm, bm)
        return out

    def viterbi(self, trans):
        tm = self.mat(trans)
        T = trans.shape[0]
        path = np.zeros(T + 1, np.int32); qpath = np.zeros(T + 1, np.float32)
        score = self.lib.decode_crf_flipflop(tm, False, path.ctypes.data_as(POINTER(c_int)), _fp(qpath))
        self.free(tm)
        return score, path, qpath

    def transpost(self, trans, return_log=True):
        tm = self.mat(trans)
        out = self.dense(self.lib.transpost_crf_flipflop(tm, return_log))
        self.free(tm)
        return out

    def trace(self, tpost_prob):
        tm = self.mat(tpost_prob)
        T, nr = tpost_prob.shape
        nbase = int(round((-1.0 + np.sqrt(1.0 + 2.0 * nr)) / 2.0))
        tr = self.lib.trace_from_posterior(tm)
        out = np.zeros((T + 1, 2 * nbase), np.int32)
        self.lib.ffref_imat_to_dense(tr, out.ctypes.data_as(_i32p))
        self.lib.free_flappie_imatrix(tr)
        self.free(tm)
        return out

    def quantile(self, x, p):
        """reference quantilef (src/util.c:100-137), one quantile"""
        x = np.ascontiguousarray(x, np.float32)
        q = ctypes.c_float(p)
        self.lib.quantilef(_fp(x), ctypes.c_size_t(x.shape[0]), ctypes.byref(q), ctypes.c_size_t(1))
        return np.float32(q.value)

    def mad(self, x):
        """reference madf (src/util.c:163-187) with the median computed inside"""
        x = np.ascontiguousarray(x, np.float32)
        self.lib.madf.restype = ctypes.c_float
        return np.float32(self.lib.madf(_fp(x), ctypes.c_size_t(x.shape[0]), None))

    # -- run-length head: the reference's own object code -------------------------
    def globalnorm_runlength(self, h, W, b, temperature=1.0):
        L = self.lib
        L.globalnorm_runlengthV2.restype = self.P
        L.globalnorm_runlengthV2.argtypes = [self.P, self.P, self.P, c_float, self.P]
        hm, Wm, bm = self.mat(h), self.mat(W), self.mat(np.asarray(b, np.float32).reshape(1, -1))
        out = L.globalnorm_runlengthV2(hm, Wm, bm, temperature, None)
        res = self.dense(out)
        self.free(hm, Wm, bm)
        return res

    def rle_viterbi(self, param):
        L = self.lib
        L.decode_crf_runlength.restype = c_float
        L.decode_crf_runlength.argtypes = [self.P, POINTER(c_int)]
        T = param.shape[0]
        pm = self.mat(param)
        path = np.zeros(T + 2, np.int32)
        score = L.decode_crf_runlength(pm, path.ctypes.data_as(POINTER(c_int)))
        self.free(pm)
        return score, path[:T]

    def rle_transpost(self, param):
        L = self.lib
        L.transpost_crf_runlength.restype = self.P
        L.transpost_crf_runlength.argtypes = [self.P]
        pm = self.mat(param)
        out = L.transpost_crf_runlength(pm)
        res = self.dense(out)
        self.free(pm)
        return res

    def format_record(self, fmt, path, uuid, readname, uuid_primary, prefix, score, nblock, basecall, quality, n, start, end):
        """append one fasta / fastq / sam record to `path` with the reference's fprintf_format"""
        L = self.lib
        L.ffref_format.restype = c_int
        L.ffref_format.argtypes = [c_char_p, c_char_p, c_char_p, c_char_p, c_bool, c_char_p, c_float, c_size_t, c_char_p, c_char_p,
                                   c_size_t, c_size_t, c_size_t]
        r = L.ffref_format(fmt.encode(), path.encode(), uuid.encode(), readname.encode(), uuid_primary, prefix.encode(),
                           score, nblock, basecall.encode(), quality.encode() if quality is not None else None, n, start, end)
        assert r == 0

    def medmad_normalise(self, x):
        x = np.array(x, np.float32, copy=True)
        self.lib.medmad_normalise_array(_fp(x), x.shape[0])
        return x

    # -- whole model -------------------------------------------------------------
    def model(self, m):
        rm = _RModel()
        rm.kind = m.kind
        rm.nconv = len(m.conv_W)
        for i in range(rm.nconv):
            rm.conv_W[i] = self.conv_mat(np.asarray(m.conv_W[i], np.float32))
            rm.conv_b[i] = self.mat(np.asarray(m.conv_b[i], np.float32).reshape(1, -1))
            rm.conv_stride[i] = m.conv_stride[i]
        for i in range(5):
            rm.iW[i] = self.mat(m.iW[i]); rm.sW[i] = self.mat(m.sW[i])
            rm.b[i] = self.mat(np.asarray(m.b[i], np.float32).reshape(1, -1))
        rm.FF_W = self.mat(m.ff_W); rm.FF_b = self.mat(np.asarray(m.ff_b, np.float32).reshape(1, -1))
        return rm

    def transitions(self, rm, signal, temperature=1.0, want_layers=False):
        signal = np.ascontiguousarray(signal, np.float32)
        dump = (self.P * 8)()
        t = self.lib.ffref_transitions(ctypes.byref(rm), _fp(signal), signal.shape[0], temperature,
                                       dump if want_layers else None)
        if not t:
            return None
        trans = self.dense(t)
        if not want_layers:
            return trans
        conv = self.dense(dump[rm.nconv - 1])
        for i in range(rm.nconv - 1):
            self.free(dump[i])
        layers = [self.dense(dump[3 + i]) for i in range(5)]
        return trans, conv, layers

    def decode(self, trans, viterbi_only=False, want_trace=True):
        tm = self.mat(trans)
        T, nr = trans.shape
        nbase = int(round((-1.0 + np.sqrt(1.0 + 2.0 * nr)) / 2.0))
        path = np.zeros(T + 2, np.int32); qpath = np.zeros(T + 2, np.float32)
        bc = ctypes.create_string_buffer(T + 2); ql = ctypes.create_string_buffer(T + 2)
        score = c_float(0)
        trace = np.zeros((T + 1, 2 * nbase), np.int32)
        post = np.zeros((T, nr), np.float32)
        n = self.lib.ffref_decode(tm, viterbi_only, path.ctypes.data_as(POINTER(c_int)), _fp(qpath), bc, ql,
                                  ctypes.byref(score), trace.ctypes.data_as(_i32p) if want_trace else None, _fp(post))
        self.free(tm)
        return dict(basecall=bc.raw[:n].decode(), quality=ql.raw[:n].decode(), score=score.value,
                    path=path[:T + 1], qpath=qpath[:T + 1], trace=trace if want_trace else None, post=post, nblock=T)

    def basecall(self, rm, signal, temperature=1.0, viterbi_only=False):
        signal = np.ascontiguousarray(signal, np.float32)
        n = signal.shape[0]
        bc = ctypes.create_string_buffer(n + 2); ql = ctypes.create_string_buffer(n + 2)
        score = c_float(0)
        nb = self.lib.ffref_basecall(ctypes.byref(rm), _fp(signal), n, temperature, viterbi_only, bc, ql,
                                     ctypes.byref(score))
        if nb < 0:
            return None
        return dict(basecall=bc.raw[:nb].decode(), quality=ql.raw[:nb].decode(), score=score.value)


def have_ref() -> bool:
    return os.path.exists(REF_SO)
