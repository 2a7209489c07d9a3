"""Throughput on mixed read lengths (BASELINE configs[3]: r10C_pcr, lengths log-uniform 1k-50k) next to a uniform batch
with the same number of blocks: how well does the group schedule of the recurrent kernel keep the slots busy?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from flappie_b200 import api
from flappie_b200.model import FlipflopModel, synthetic_reads
from flappie_b200.signal import prepare_read

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1000, 50000)
fm = FlipflopModel.for_name("r10C_pcr", seed=1)
rng = np.random.default_rng(5)
lens = np.exp(rng.uniform(np.log(lo), np.log(hi), size=n)).astype(np.int64)
model = api.Model(fm, device=0)
st = torch.cuda.Stream()
ctx = api.Context(model, stream=st.cuda_stream)
for name, ls in ((f"mixed {lo}-{hi}", lens), ("uniform", np.full(n, int(lens.mean()), np.int64))):
    reads = [prepare_read(r) for r in synthetic_reads(n, ls, seed=13)]
    ll = np.array([len(r) for r in reads], np.int64)
    off = np.zeros(n + 1, np.int64); np.cumsum(ll, out=off[1:])
    b, o = ctx.make_batch(np.concatenate(reads), off, 1.0, 0)
    ctx.upload(b)
    for _ in range(2):
        ctx.forward()
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(3):
        ctx.forward()
    ctx.sync()
    ms = (time.perf_counter() - t0) * 1e3 / 3
    g = ctx.forward_timed()
    print(f"{name:13s}: {n} reads, {int(ls.sum())/1e6:.1f} M raw samples, max {int(ls.max())}: {ms:8.1f} ms/step = "
          f"{ls.sum() / ms / 1e3:7.1f} M samples/s   (rnn {g['rnn']:.1f} ms, gemm {g['gemm']:.1f} ms)")
