"""Host-side signal preparation that precedes the hot path in calculate_post
(reference src/flappie.c:251-259): trimming by MAD segmentation and med-MAD (or delta)
normalisation.  These stay on the host, as in the reference (SURVEY.md section 8(f) ranks
moving them to the device as the next step after the hot path); they are restated here in
numpy float32 so the batched driver and the benchmarks feed the GPU exactly what the
reference would feed calculate_transitions().
"""
from __future__ import annotations

import numpy as np


def quantilef(x: np.ndarray, p: float) -> np.float32:
    """reference src/util.c:100-137 (linear interpolation on the sorted array)."""
    space = np.sort(np.asarray(x, np.float32))
    nx = space.shape[0]
    pos = np.float32(p) * np.float32(nx - 1)
    idx = int(pos)
    remf = np.float32(pos - np.float32(idx))
    if idx < nx - 1:
        # C: `(1.0 - remf) * space[idx] + remf * space[idx + 1]` -- the first product is double, the second a
        # float product that is then widened
        return np.float32((1.0 - float(remf)) * float(space[idx]) + float(np.float32(remf * space[idx + 1])))
    return space[idx]


def medianf(x: np.ndarray) -> np.float32:
    return quantilef(x, 0.5)


def madf(x: np.ndarray, med=None) -> np.float32:
    """reference src/util.c:163-187"""
    x = np.asarray(x, np.float32)
    if x.shape[0] == 1:
        return np.float32(0.0)
    m = medianf(x) if med is None else np.float32(med)
    return np.float32(medianf(np.abs(x - m)) * np.float32(1.4826))


def medmad_normalise_array(x: np.ndarray) -> np.ndarray:
    """reference src/util.c:198-212"""
    x = np.array(x, np.float32, copy=True)
    if x.shape[0] == 1:
        x[0] = 0.0
        return x
    xmed = medianf(x)
    xmad = madf(x, xmed)
    return ((x - xmed) / xmad).astype(np.float32)


def difference_array(x: np.ndarray) -> np.ndarray:
    """reference src/util.c:278-287"""
    x = np.array(x, np.float32, copy=True)
    x[:-1] = x[1:] - x[:-1]
    x[-1] = 0.0
    return x


def trim_raw_by_mad(raw: np.ndarray, start: int, end: int, chunk_size: int, perc: float):
    """reference src/flappie_common.c:47-81 -> (start, end)"""
    nsample = end - start
    nchunk = nsample // chunk_size
    end = nchunk * chunk_size
    if nchunk == 0:
        return start, end
    seg = np.asarray(raw[start:start + nchunk * chunk_size], np.float32).reshape(nchunk, chunk_size)
    mad = np.array([madf(c) for c in seg], np.float32)
    thresh = quantilef(mad, perc)
    for i in range(nchunk):
        if mad[i] > thresh:
            break
        start += chunk_size
    for i in range(nchunk, 0, -1):
        if mad[i - 1] > thresh:
            break
        end -= chunk_size
    return start, end


def trim_and_segment_raw(raw: np.ndarray, trim_start: int = 200, trim_end: int = 10, varseg_chunk: int = 100,
                         varseg_thresh: float = 0.0):
    """reference src/flappie_common.c:13-28 with the CLI defaults of src/flappie.c:105-107.
    Returns (start, end) or None if nothing is left."""
    n = raw.shape[0]
    start, end = trim_raw_by_mad(raw, 0, n, varseg_chunk, varseg_thresh)
    start = start + trim_start if (n - start) > trim_start else n
    end = end - trim_end if end > trim_end else 0
    if start >= end:
        return None
    return start, end


def prepare_read(raw: np.ndarray, delta: float = 0.0, trim=(200, 10), segmentation=(100, 0.0)):
    """calculate_post up to the network call (src/flappie.c:248-259): returns the normalised
    trimmed signal (float32) or None."""
    se = trim_and_segment_raw(raw, trim[0], trim[1], segmentation[0], segmentation[1])
    if se is None:
        return None
    x = np.asarray(raw[se[0]:se[1]], np.float32)
    if delta == 0.0:
        return medmad_normalise_array(x)
    return (difference_array(x) / np.float32(delta)).astype(np.float32)
