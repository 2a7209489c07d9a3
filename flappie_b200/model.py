"""Flip-flop model weight bundles in the reference's own layout.

The reference compiles its weights in as generated C headers holding `_Mat`
structs (reference src/flappie_matrix.h:18-24; format written by
misc/taiyaki_flipflop5_guppy.py:38-99) and binds them into `guppy_model` /
`guppy_stride5_model` bundles (reference src/networks.c:150-215).  Those headers
are git-LFS pointers in the checkout, so this module provides

* `FlipflopModel`  -- dense numpy weights in the bundle's field order,
* `FlipflopModel.synthetic()` -- a seeded generator at the model sizes of
  SURVEY.md section 8(a-0), used by the tests and by bench.py,
* `to_mat_bundle()` -- the exact padded `_Mat` memory images (column-major, rows
  padded to a multiple of 4 floats, convolution rows `nf4*winlen - nf4 + nf`)
  that the C-ABI `ffb_model_create()` accepts, i.e. what a maintainer would pass
  straight out of the compiled `.mdl` headers,
* `save()` / `load()` -- a flat `.npz` dump.
"""
from __future__ import annotations

import ctypes
import dataclasses
from typing import List, Tuple

import numpy as np

KIND_GRU = 0   # guppy_model           (networks.c:150-177)  conv(tanh) + 5 grumod
KIND_LSTM = 1  # guppy_stride5_model   (networks.c:180-215)  3 conv(swish) + 5 lstm

# name -> (kind, size, nbase) following networks.c:21-40 plus the flappie-1.x
# alias the north star names.
MODEL_TABLE = {
    "r941_native": (KIND_LSTM, 384, 4),
    "r941_rna002": (KIND_LSTM, 256, 4),
    "r941_5mC": (KIND_GRU, 256, 5),
    "r103_native": (KIND_LSTM, 512, 4),
    "r10C_pcr": (KIND_GRU, 256, 4),
    "r941_native_gru": (KIND_GRU, 256, 4),   # north-star topology of r941_native (flappie 1.x)
    "rle_r941_native": (KIND_LSTM, 256, 4),  # runnie: LSTM stack + run-length head (size inferred like the others)
}


class Mat(ctypes.Structure):
    """Layout-compatible with the reference `_Mat` (flappie_matrix.h:18-24)."""

    _fields_ = [
        ("nr", ctypes.c_size_t),
        ("nrq", ctypes.c_size_t),
        ("nc", ctypes.c_size_t),
        ("stride", ctypes.c_size_t),
        ("data", ctypes.POINTER(ctypes.c_float)),
    ]


def pad4(n: int) -> int:
    return 4 * ((n + 3) // 4)


def mat_image(dense_cols: np.ndarray, nr: int | None = None) -> Tuple[Mat, np.ndarray]:
    """Padded `_Mat` image of a matrix given as [nc][rows] (one row per flappie column)."""
    dense_cols = np.ascontiguousarray(dense_cols, dtype=np.float32)
    nc, rows = dense_cols.shape
    if nr is None:
        nr = rows
    stride = pad4(nr)
    buf = np.zeros((nc, stride), dtype=np.float32)
    buf[:, :rows] = dense_cols
    m = Mat(nr, stride // 4, nc, stride, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    return m, buf


def conv_mat_image(W: np.ndarray) -> Tuple[Mat, np.ndarray]:
    """W [nfilter][winlen][nf] -> reference convolution filter `_Mat`
    (misc/taiyaki_flipflop5_guppy.py:86-95: features padded to 4 per tap,
    nr = nf4*winlen - nf4 + nf)."""
    nfilter, winlen, nf = W.shape
    nf4 = pad4(nf)
    nr = nf4 * winlen - nf4 + nf
    stride = pad4(nr)
    buf = np.zeros((nfilter, stride), dtype=np.float32)
    for k in range(winlen):
        buf[:, k * nf4:k * nf4 + nf] = W[:, k, :]
    m = Mat(nr, stride // 4, nfilter, stride, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    return m, buf


@dataclasses.dataclass
class FlipflopModel:
    kind: int
    conv_W: List[np.ndarray]      # each [nfilter][winlen][nf]
    conv_b: List[np.ndarray]      # each [nfilter]
    conv_stride: List[int]
    iW: List[np.ndarray]          # 5 x [G*S][in]
    sW: List[np.ndarray]          # 5 x [G*S][S]
    b: List[np.ndarray]           # 5 x [G*S]
    ff_W: np.ndarray              # [nparam][S]
    ff_b: np.ndarray              # [nparam]
    name: str = "synthetic"
    head: str = "flipflop"        # "runlength": the LSTM stack with runnie's head (networks.c:675-722)

    @property
    def size(self) -> int:
        return self.sW[0].shape[1]

    @property
    def ngate(self) -> int:
        return 3 if self.kind == KIND_GRU else 4

    @property
    def nparam(self) -> int:
        return self.ff_W.shape[0]

    @property
    def nbase(self) -> int:
        return int(round((-1.0 + np.sqrt(1.0 + 2.0 * self.nparam)) / 2.0))

    @property
    def nstate(self) -> int:
        return 2 * self.nbase

    @property
    def stride(self) -> int:
        s = 1
        for x in self.conv_stride:
            s *= x
        return s

    def nblock(self, nsample: int) -> int:
        """Blocks produced for `nsample` samples: iceil per conv (layers.c:204)."""
        t = nsample
        for w, s in zip(self.conv_W, self.conv_stride):
            if t < w.shape[1]:
                return -1
            t = (t + s - 1) // s
        return t

    # ------------------------------------------------------------------ synthetic
    @staticmethod
    def synthetic(kind: int = KIND_GRU, size: int = 256, nbase: int = 4, seed: int = 1,
                  conv_stride: int | None = None, name: str | None = None,
                  winlen: int = 19, saturate: bool = False) -> "FlipflopModel":
        """Seeded weights at a reference model's shape.
        Scales keep the gates out of saturation and the recurrence contractive, as for
        trained models.  saturate=True: gate biases drawn from U(-4, 4) instead of U(-0.1, 0.1),
        so that a good part of the sigmoids / tanhs sit pinned at 0 / 1 / +-1 for whole reads (the
        wide-dynamic-range regime of trained gates) while the gains -- and with them the
        layer-to-layer amplification of rounding noise -- stay what they are."""
        rng = np.random.default_rng(seed)
        G = 3 if kind == KIND_GRU else 4

        def u(shape, a):
            return rng.uniform(-a, a, size=shape).astype(np.float32)

        conv_W, conv_b, strides = [], [], []
        if kind == KIND_GRU:
            s = 2 if conv_stride is None else conv_stride
            conv_W.append(u((size, winlen, 1), 0.35))
            conv_b.append(u((size,), 0.2))
            strides.append(s)
            g_i, g_s, g_f = 3.0, 1.5, 6.0
        else:
            s = 5 if conv_stride is None else conv_stride
            conv_W += [u((4, 5, 1), 0.8), u((16, 5, 4), 0.6), u((size, winlen, 16), 0.25)]
            conv_b += [u((4,), 0.2), u((16,), 0.2), u((size,), 0.2)]
            strides += [1, 1, s]
            g_i, g_s, g_f = 5.0, 2.0, 8.0
        # gains chosen (tests/golden/README.md) so that the output follows the signal
        # (hundreds of base transitions per read) while a 1e-6 input perturbation grows
        # by < 10x through the five layers, like a trained, contractive model
        iW = [u((G * size, size), g_i / np.sqrt(size)) for _ in range(5)]
        sW = [u((G * size, size), g_s / np.sqrt(size)) for _ in range(5)]
        b = [u((G * size,), 4.0 if saturate else 0.1) for _ in range(5)]
        nstate = 2 * nbase
        nparam = nstate * (nbase + 1)
        ff_W = u((nparam, size), g_f / np.sqrt(size))
        ff_b = u((nparam,), 0.1)
        return FlipflopModel(kind, conv_W, conv_b, strides, iW, sW, b, ff_W, ff_b,
                             name or f"synthetic_{'gru' if kind == KIND_GRU else 'lstm'}{size}_{nbase}b_s{seed}")

    @staticmethod
    def for_name(model: str, seed: int = 1) -> "FlipflopModel":
        kind, size, nbase = MODEL_TABLE[model]
        fm = FlipflopModel.synthetic(kind, size, nbase, seed, name=model)
        if model.startswith("rle_"):
            fm.head = "runlength"
        return fm

    # ------------------------------------------------------------------ I/O
    def save(self, path: str) -> None:
        d = {"kind": np.int32(self.kind), "conv_stride": np.asarray(self.conv_stride, np.int32),
             "ff_W": self.ff_W, "ff_b": self.ff_b, "name": np.str_(self.name)}
        for i, (w, bb) in enumerate(zip(self.conv_W, self.conv_b)):
            d[f"conv{i}_W"], d[f"conv{i}_b"] = w, bb
        for i in range(5):
            d[f"l{i}_iW"], d[f"l{i}_sW"], d[f"l{i}_b"] = self.iW[i], self.sW[i], self.b[i]
        np.savez(path, **d)

    @staticmethod
    def load(path: str) -> "FlipflopModel":
        z = np.load(path)
        strides = [int(x) for x in z["conv_stride"]]
        n = len(strides)
        return FlipflopModel(int(z["kind"]), [z[f"conv{i}_W"] for i in range(n)],
                             [z[f"conv{i}_b"] for i in range(n)], strides,
                             [z[f"l{i}_iW"] for i in range(5)], [z[f"l{i}_sW"] for i in range(5)],
                             [z[f"l{i}_b"] for i in range(5)], z["ff_W"], z["ff_b"], str(z["name"]))

    def save_bundle(self, path: str) -> None:
        """Binary weight bundle for the C command line (flappie_b200/host/ffb_host.h): the `_Mat` images of
        to_mat_bundle() in the reference's struct order, padded columns included."""
        import struct
        mats, keep = self.to_mat_bundle()
        strides = list(self.conv_stride) + [0] * (3 - len(self.conv_stride))
        with open(path, "wb") as fh:
            fh.write(b"FFBW1\0\0\0")
            kind = 2 if self.head == "runlength" else self.kind      # FFB_KIND_RUNLENGTH
            fh.write(struct.pack("<6i", kind, len(self.conv_stride), strides[0], strides[1], strides[2], len(mats)))
            for m in mats:
                fh.write(struct.pack("<2Q", m.nr, m.nc))
                fh.write(np.ctypeslib.as_array(m.data, shape=(m.nc * m.stride,)).astype("<f4").tobytes())
        del keep

    # ------------------------------------------------------------------ `_Mat` bundle
    def to_mat_bundle(self):
        """Return (mats, keepalive): `mats` is the list of `_Mat` in the field order of
        guppy_model (networks.c:150-177: conv_W, conv_b, then iW,sW,b x5, FF_W, FF_b) or
        guppy_stride5_model (networks.c:180-215: conv{1,2,3}_{W,b}, then the same)."""
        mats, keep = [], []
        for w, bb in zip(self.conv_W, self.conv_b):
            m, buf = conv_mat_image(w)
            mats.append(m); keep.append(buf)
            m, buf = mat_image(bb.reshape(1, -1))
            mats.append(m); keep.append(buf)
        for i in range(5):
            for arr in (self.iW[i], self.sW[i]):
                m, buf = mat_image(arr)
                mats.append(m); keep.append(buf)
            m, buf = mat_image(self.b[i].reshape(1, -1))
            mats.append(m); keep.append(buf)
        m, buf = mat_image(self.ff_W)
        mats.append(m); keep.append(buf)
        m, buf = mat_image(self.ff_b.reshape(1, -1))
        mats.append(m); keep.append(buf)
        return mats, keep


def synthetic_reads(n_reads: int, n_samples, seed: int = 7) -> List[np.ndarray]:
    """pA-like squiggles: piecewise-constant levels N(90, 12^2) with
    geometric dwell (mean 9 samples) plus N(0, 1.5^2) noise.  `n_samples` is an int or
    a per-read sequence."""
    rng = np.random.default_rng(seed)
    if np.isscalar(n_samples):
        n_samples = [int(n_samples)] * n_reads
    out = []
    for n in n_samples:
        n = int(n)
        nseg = max(4, int(n / 9 * 1.5) + 8)
        dwell = rng.geometric(1.0 / 9.0, size=nseg)
        while dwell.sum() < n:
            dwell = np.concatenate([dwell, rng.geometric(1.0 / 9.0, size=nseg)])
        levels = rng.normal(90.0, 12.0, size=dwell.shape[0]).astype(np.float32)
        sig = np.repeat(levels, dwell)[:n] + rng.normal(0.0, 1.5, size=n).astype(np.float32)
        out.append(sig.astype(np.float32))
    return out
