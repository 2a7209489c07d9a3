"""Probe: does keeping two batches in flight (two contexts, two streams) raise throughput?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from flappie_b200 import api
from bench import make_workload, MODEL_CHOICES

name = sys.argv[1] if len(sys.argv) > 1 else "r941_native_gru"
nctx = int(sys.argv[2]) if len(sys.argv) > 2 else 2
steps = 10
fm, reads = make_workload(MODEL_CHOICES[name][0], 1024, 4000, seed=7)
lens = np.array([len(r) for r in reads], np.int64)
off = np.zeros(len(reads) + 1, np.int64); np.cumsum(lens, out=off[1:])
sig = np.concatenate(reads)
model = api.Model(fm, device=0)
ctxs, batches = [], []
for i in range(nctx):
    st = torch.cuda.Stream()
    c = api.Context(model, stream=st.cuda_stream)
    b, o = c.make_batch(sig, off, 1.0, 0, None)
    c.upload(b)
    ctxs.append((c, st)); batches.append((b, o))
for rep in range(3):
    for c, _ in ctxs:
        c.forward()
torch.cuda.synchronize()
for mode in ("one", "alternate"):
    t0 = time.perf_counter()
    for i in range(steps):
        ctxs[i % nctx if mode == "alternate" else 0][0].forward()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    print(f"{name} {mode:9s} ({nctx} contexts): {ms:.2f} ms/step  {1024 * 4000 / ms / 1e3:.1f} M samples/s")
