"""Slots per cluster x clusters of the recurrent kernel (csrc/api.cu:plan_groups) on the batch shapes where the choice
matters: which (G, clusters) does the cost model pick, and what do the alternatives cost?  -> profiles/r02_slots_ab.txt
    python tools/slots_ab.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from flappie_b200.api import Context, Model
from flappie_b200.model import FlipflopModel, synthetic_reads
from flappie_b200.signal import prepare_read

def mixed(n, lo, hi, seed):
    rng = np.random.default_rng(seed)
    return [int(x) for x in np.exp(rng.uniform(np.log(lo), np.log(hi), n))]

shapes = [("1024 x 4000 (configs[1])", "r941_native_gru", [4000] * 1024),
          ("4096 x 4000 (configs[2])", "r941_5mC", [4000] * 4096),
          ("2048 reads log-uniform 1k-50k (configs[3])", "r10C_pcr", mixed(2048, 1000, 50000, 100)),
          ("4096 reads log-uniform 1k-10k", "r10C_pcr", mixed(4096, 1000, 10000, 101))]
print(f"# {time.strftime('%Y-%m-%d %H:%M')}: ms per ffb_forward (median of 3, forward-backward mode), streamed schedule; 'auto' = the planner's own choice")
for title, name, lens in shapes:
    fm = FlipflopModel.for_name(name, seed=1)
    raws = synthetic_reads(len(lens), lens, seed=7)
    m = Model(fm); ctx = Context(m)
    lensa = np.array([len(r) for r in raws], np.int64)
    off = np.zeros(len(raws) + 1, np.int64); np.cumsum(lensa, out=off[1:])
    raw = np.concatenate(raws)
    del raws
    import ctypes, torch
    samples = int(lensa.sum())
    for tag, env in (("auto", {}), ("auto, GEMM not streamed", {"FFB_NO_STREAM_GEMM": "1"}), ("G=5 x13", {"FFB_TC_SLOTS": "5", "FFB_TC_CLUSTERS": "13"}), ("G=5 x15", {"FFB_TC_SLOTS": "5", "FFB_TC_CLUSTERS": "15"}),
                     ("G=6 x13", {"FFB_TC_SLOTS": "6", "FFB_TC_CLUSTERS": "13"}), ("G=6 x15", {"FFB_TC_SLOTS": "6", "FFB_TC_CLUSTERS": "15"}),
                     ("G=4 x15", {"FFB_TC_SLOTS": "4", "FFB_TC_CLUSTERS": "15"}), ("G=3 x15", {"FFB_TC_SLOTS": "3", "FFB_TC_CLUSTERS": "15"}),
                     ("G=2 x15", {"FFB_TC_SLOTS": "2", "FFB_TC_CLUSTERS": "15"}), ("G=1 x15", {"FFB_TC_SLOTS": "1", "FFB_TC_CLUSTERS": "15"})):
        for k in ("FFB_TC_SLOTS", "FFB_TC_CLUSTERS", "FFB_NO_STREAM_GEMM"):
            os.environ.pop(k, None)
        os.environ.update(env)
        b, o = ctx.make_batch(raw, off, 1.0, 0, want_path=True)
        rb, st, en = ctx.make_raw_batch(raw, off)
        ctx._check(ctx.lib.lib.ffb_upload_raw(ctx.handle, ctypes.byref(rb), ctypes.byref(b)), "upload")
        ctx.forward(); ctx.sync()
        ts = []
        for _ in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter(); ctx.forward(); ctx.sync(); ts.append((time.perf_counter() - t0) * 1e3)
        g = ctx.forward_timed()
        print(f"{title:45s} {tag:24s} {np.median(ts):8.2f} ms  {samples / np.median(ts) / 1e3:7.1f} M samples/s   sequential: rnn {g['rnn']:.1f} gemm {g['gemm']:.1f} ms", flush=True)
    ctx.close(); m.close()
