"""Saturated-gate weight seeds (FlipflopModel.synthetic(saturate=True)) through the tensor path and the fp32 CUDA-core path,
against the oracle: max |d trans| of each.  Used to calibrate FFB_ACC_COMP (profiles/r02_acc_comp.txt)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from flappie_b200 import signal as hs
from flappie_b200.api import Context, Model
from flappie_b200.model import synthetic_reads
from test_gpu_hardening import _model, _pool_map

for name, sat in (("lstm256", True), ("gru256", True), ("lstm384", True), ("lstm256", False), ("gru256", False)):
    fm = _model(name, 3, sat=sat)
    sigs = [hs.prepare_read(r) for r in synthetic_reads(24, 2500, seed=13)]
    m = Model(fm); ctx = Context(m)
    simt = ctx.basecall(sigs, viterbi_only=True, want_trans=True, fp32_simt=True)
    res = ctx.basecall(sigs, viterbi_only=True, want_trans=True)
    outs = _pool_map([((name, 3, sat), s, True) for s in sigs])
    dt = max(float(np.max(np.abs(res.read_trans(i) - o["trans"]))) for i, o in enumerate(outs))
    ds = max(float(np.max(np.abs(simt.read_trans(i) - o["trans"]))) for i, o in enumerate(outs))
    mean_t = float(np.mean([np.mean(res.read_trans(i) - o["trans"]) for i, o in enumerate(outs)]))
    nd = sum(int(np.count_nonzero(res.read_path(i)[0] != o["vit_path"])) for i, o in enumerate(outs))
    print(f"{name:8s} saturate={int(sat)}: max|d trans| tensor {dt:.2e}  fp32 CUDA-core {ds:.2e}   mean signed d (tensor) {mean_t:+.2e}   differing Viterbi blocks {nd}", flush=True)
    ctx.close(); m.close()
