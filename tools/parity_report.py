"""profiles/r02_parity_report.txt: the CUDA path against the oracle at BASELINE sizes, as measured numbers.

    python tools/parity_report.py > profiles/r02_parity_report.txt       (GPU box; ~2 min of host time for the oracle)

For each model: N reads of the bench workload (4000 raw samples) through the C ABI in --viterbi and default mode, against
oracle/flappie_oracle.c on the same normalised signal: max |d trans|, reads whose Viterbi path differs anywhere, reads whose
called bases differ anywhere (default mode), bases compared.  Then the long reads of configs[3]."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from flappie_b200 import signal as hs
from flappie_b200.api import Context, Library, Model
from flappie_b200.model import synthetic_reads
from test_gpu_hardening import _model, _pool_map

lib = Library.get()
print(f"# parity report, {time.strftime('%Y-%m-%d %H:%M:%S')}, {lib.lib.ffb_version().decode()}, oracle = oracle/flappie_oracle.c (scalar fp32 restatement,")
print("# pinned to the reference's object code by tests/test_oracle.py); tolerance of north_star: 1e-4 on trans, Viterbi path bit-exact")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for name, seed in (("gru256", 1), ("lstm384", 1), ("lstm256", 1), ("gru256_5", 1)):
    fm = _model(name, seed)
    sigs = [hs.prepare_read(r) for r in synthetic_reads(n, 4000, seed=7)]
    m = Model(fm); ctx = Context(m)
    rv = ctx.basecall(sigs, viterbi_only=True, want_trans=True)
    rf = ctx.basecall(sigs, viterbi_only=False)
    t0 = time.time()
    outs = _pool_map([((name, seed, False), s, False) for s in sigs])
    dmax = 0.0; bad_v = bad_f = nbase_tot = nbase_diff = 0
    for i, o in enumerate(outs):
        dmax = max(dmax, float(np.max(np.abs(rv.read_trans(i) - o["trans"]))))
        bad_v += int(not np.array_equal(rv.read_path(i)[0], o["vit_path"]))
        bases, quals = lib.emit_bases(*rf.read_path(i), fm.nbase)
        want = o["fb"]["basecall"]
        nbase_tot += len(want)
        if bases != want:
            bad_f += 1
            nbase_diff += sum(a != b for a, b in zip(bases, want)) + abs(len(bases) - len(want))
    print(f"{name:9s} {n} reads x 4000 samples: max|d trans| = {dmax:.3e}   reads with a differing Viterbi path: {bad_v}/{n}   "
          f"default mode: reads with a differing base: {bad_f}/{n} ({nbase_diff} of {nbase_tot} bases)   [oracle {time.time() - t0:.0f} s]", flush=True)
    ctx.close(); m.close()
# long reads (configs[3]: 1 k - 50 k samples)
for name, lens in (("gru256", [50000, 35000, 20000]), ("lstm384", [20000]), ("lstm256", [30000])):
    fm = _model(name, 1)
    raws = synthetic_reads(len(lens) + 5, lens + [1000, 2300, 7000, 12000, 3100], seed=5)
    sigs = [hs.prepare_read(r) for r in raws]
    m = Model(fm); ctx = Context(m)
    rv = ctx.basecall(sigs, viterbi_only=True, want_trans=True)
    outs = _pool_map([((name, 1, False), sigs[i], True) for i in range(len(lens))])
    for i, o in enumerate(outs):
        d = float(np.max(np.abs(rv.read_trans(i) - o["trans"])))
        p = rv.read_path(i)[0]
        print(f"{name:9s} long read {lens[i]:6d} samples = {len(p) - 1:6d} steps: max|d trans| = {d:.3e}   differing Viterbi blocks: "
              f"{int(np.count_nonzero(p != o['vit_path']))}/{len(p)}", flush=True)
    ctx.close(); m.close()
