"""Timing of the recurrent layers of configs[1] (1024 x 4000 samples, GRU-256) for a library built with
FFB_EXTRA_NVCC_FLAGS=-DFFB_RNN_ABLATE=<mask> (csrc/rnn_tc.cu): what does each part of the step cost the layer?
Results of such a build are garbage; only the time is read.   python tools/ablate_timing.py <tag>"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from flappie_b200.api import Context, Model
from flappie_b200.model import FlipflopModel, synthetic_reads

tag = sys.argv[1] if len(sys.argv) > 1 else "?"
fm = FlipflopModel.for_name("r941_native_gru", seed=1)
raws = synthetic_reads(1024, [4000] * 1024, seed=7)
m = Model(fm); ctx = Context(m)
lensa = np.array([len(r) for r in raws], np.int64)
off = np.zeros(len(raws) + 1, np.int64); np.cumsum(lensa, out=off[1:])
raw = np.concatenate(raws)
for cfg, env in (("5x13", {}), ("3x15", {"FFB_TC_SLOTS": "3", "FFB_TC_CLUSTERS": "15"}), ("1x15", {"FFB_TC_SLOTS": "1", "FFB_TC_CLUSTERS": "15"})):
    for k in ("FFB_TC_SLOTS", "FFB_TC_CLUSTERS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    b, o = ctx.make_batch(raw, off, 1.0, 0, want_path=True)
    rb, st, en = ctx.make_raw_batch(raw, off)
    ctx._check(ctx.lib.lib.ffb_upload_raw(ctx.handle, ctypes.byref(rb), ctypes.byref(b)), "upload")
    ctx.forward(); ctx.sync()
    g = [ctx.forward_timed() for _ in range(3)]
    rnn = float(np.median([x["rnn"] for x in g]))
    rounds = {"5x13": 1, "3x15": 2, "1x15": 5}[cfg]
    print(f"{tag:10s} {cfg}: rnn {rnn:6.2f} ms = {rnn * 1e3 / (5 * 1895 * rounds):.3f} us per step and round   gemm {np.median([x['gemm'] for x in g]):.2f}", flush=True)
ctx.close(); m.close()
