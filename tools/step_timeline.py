"""Where does a step go?  Timeline of two consecutive steps of configs[1] from globaltimer stamps written by the recurrent
kernel and the W-stationary GEMM themselves (entry / exit of CTA 0, and of the GEMM's last CTA) -- no profiler in the way, the
streamed schedule as it runs.  Needs FFB_EXTRA_NVCC_FLAGS=-DFFB_RNN_PROFILE.    python tools/step_timeline.py [model] [reads]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from flappie_b200.api import Context, Library, Model
from flappie_b200.model import FlipflopModel, synthetic_reads
from flappie_b200.signal import prepare_read

name = sys.argv[1] if len(sys.argv) > 1 else "r941_native_gru"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
L = Library.get().lib
L.ffb_test_timeline.restype = ctypes.c_int
L.ffb_test_timeline.argtypes = [ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_int), ctypes.c_int]
fm = FlipflopModel.for_name(name, seed=1)
reads = [prepare_read(r) for r in synthetic_reads(n, 4000, seed=7)]
lens = np.array([len(r) for r in reads], np.int64)
off = np.zeros(n + 1, np.int64); np.cumsum(lens, out=off[1:])
sig = np.concatenate(reads)
m = Model(fm); ctx = Context(m)
b, o = ctx.make_batch(sig, off, 1.0, 0, want_path=True)
ctx.upload(b)
for _ in range(3):
    ctx.forward()
ctx.sync()
r = (ctypes.c_uint64 * 64)(); g = (ctypes.c_uint64 * 128)(); cnt = (ctypes.c_int * 2)()
assert L.ffb_test_timeline(r, g, cnt, 1) == 1, "library not built with -DFFB_RNN_PROFILE"
import time
t0 = time.perf_counter()
for _ in range(2):
    ctx.forward()
ctx.sync()
wall = (time.perf_counter() - t0) * 1e3
L.ffb_test_timeline(r, g, cnt, 0)
ev = []
for k in range(min(cnt[0], 32)):
    ev.append((int(r[2 * k]), int(r[2 * k + 1]), f"rnn  launch {k}"))
for k in range(min(cnt[1], 32)):
    ev.append((int(g[4 * k]), max(int(g[4 * k + 1]), int(g[4 * k + 3])), f"gemm launch {k}  (CTA 0 {int(g[4*k+1]) - int(g[4*k])} ns; last CTA enters +{int(g[4*k+2]) - int(g[4*k])} ns, leaves +{int(g[4*k+3]) - int(g[4*k])} ns)"))
ev.sort()
base = ev[0][0]
print(f"# {name}, {n} reads x 4000 samples, two steps back to back on one context: wall {wall:.2f} ms; times in us from the first stamp")
prev_end = None
busy_rnn = 0.0
for s, e, what in ev:
    gap = "" if prev_end is None else f"   [{(s - prev_end) / 1e3:+8.1f} us after the previous kernel's end]"
    print(f"{(s - base) / 1e3:10.1f} .. {(e - base) / 1e3:10.1f}  ({(e - s) / 1e3:8.1f} us)  {what}{gap}")
    prev_end = e if prev_end is None else max(prev_end, e)
    if what.startswith("rnn"): busy_rnn += (e - s) / 1e3
print(f"# recurrent kernels: {busy_rnn / 1e3:.2f} ms of {(ev[-1][1] - base) / 1e6:.2f} ms between the first and the last stamp")
L.ffb_debug_group_times.restype = ctypes.c_int
L.ffb_debug_group_times.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float)]
gt = (ctypes.c_float * 4)()
if L.ffb_debug_group_times(ctx.handle, gt) == 0:
    print(f"# CUDA events of the last step (same schedule): convolution {gt[0]:.3f} ms, layer-1 GEMM .. last recurrent layer {gt[1]:.3f} ms, output layer {gt[2]:.3f} ms, decoding {gt[3]:.3f} ms")
