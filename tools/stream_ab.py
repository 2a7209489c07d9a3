"""Streamed vs unstreamed input GEMM per model (1024 x 4000-sample reads): ms per ffb_forward, median of 5.
    python tools/stream_ab.py [model ...]"""
import os, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from flappie_b200.api import Context, Model
from flappie_b200.model import FlipflopModel, synthetic_reads

for name in (sys.argv[1:] or ["r941_native", "r941_native_gru", "r941_rna002"]):
    fm = FlipflopModel.for_name(name, seed=1)
    raws = synthetic_reads(1024, 4000, seed=7)
    off = np.zeros(1025, np.int64); np.cumsum([len(r) for r in raws], out=off[1:])
    raw = np.concatenate(raws)
    m = Model(fm); ctx = Context(m)
    for tag, env in (("streamed", {}), ("not streamed", {"FFB_NO_STREAM_GEMM": "1"}), ("streamed", {}), ("not streamed", {"FFB_NO_STREAM_GEMM": "1"})):
        os.environ.pop("FFB_NO_STREAM_GEMM", None); os.environ.update(env)
        b, o = ctx.make_batch(raw, off, 1.0, 0)
        rb, st, en = ctx.make_raw_batch(raw, off)
        ctx._check(ctx.lib.lib.ffb_upload_raw(ctx.handle, ctypes.byref(rb), ctypes.byref(b)), "upload")
        for _ in range(2):
            ctx.forward()
        ctx.sync()
        ts = []
        for _ in range(5):
            torch.cuda.synchronize(); t0 = time.perf_counter(); ctx.forward(); ctx.sync(); ts.append((time.perf_counter() - t0) * 1e3)
        print(f"{name:18s} {tag:13s} {np.median(ts):7.2f} ms  (min {min(ts):.2f})", flush=True)
    ctx.close(); m.close()
