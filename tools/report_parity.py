"""Print (not assert) the deviation of every stage of the GPU path from the oracle for the
tcgen05 shapes -- used on the GPU box when tuning the recurrent kernel's numerics.
    python tools/report_parity.py [gru|lstm]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from flappie_b200.api import Context, Model
from flappie_b200.model import KIND_GRU, KIND_LSTM, FlipflopModel
from ffb_testutil import norm_reads
from oracle.pyoracle import Oracle

kind = KIND_LSTM if (len(sys.argv) > 1 and sys.argv[1] == "lstm") else KIND_GRU
orc = Oracle()
fm = FlipflopModel.synthetic(kind, 256, 4, seed=11)
reads = norm_reads(24, 4000, seed=3)
m = Model(fm); ctx = Context(m)
for simt in (False, True):
    res = ctx.basecall(reads, viterbi_only=True, want_trans=True, keep_layers=True, fp32_simt=simt)
    layers_g = [ctx.fetch_layer(1 + l) for l in range(5)]
    dl = np.zeros(5); dt = 0.0; same = 0; nb = 0
    for i, sig in enumerate(reads):
        trans_o, conv_o, layers_o = orc.transitions(fm, sig, 1.0, want_layers=True)
        b0, b1 = int(res.blk_off[i]), int(res.blk_off[i + 1])
        for l in range(5):
            dl[l] = max(dl[l], float(np.max(np.abs(layers_g[l][b0:b1] - layers_o[l]))))
        dt = max(dt, float(np.max(np.abs(res.read_trans(i) - trans_o))))
        _, p_o, _ = orc.viterbi(trans_o)
        p_g, _ = res.read_path(i)
        same += int(np.array_equal(p_g, p_o)); nb += 1
    if not simt:
        res2 = ctx.basecall(reads, viterbi_only=True, want_trans=True)     # no keep_layers: the streamed-GEMM schedule
        d2 = max(float(np.max(np.abs(res2.read_trans(i) - res.read_trans(i)))) for i in range(len(reads)))
        print(f"streamed vs sequential schedule: max|d trans| = {d2:.3e} (same arithmetic: expect 0)", flush=True)
    print(f"{'fp32 SIMT' if simt else 'tensor   '}: layer max|d| {' '.join(f'{x:.2e}' for x in dl)}  trans {dt:.2e}  identical Viterbi paths {same}/{nb}", flush=True)
