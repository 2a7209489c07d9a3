"""Parity where it used to be thin ("what's weak" 1-4): the LONGEST reads against the oracle, measured
base-for-base mismatch rates over hundreds of reads instead of an `n - 1 of n` allowance, saturated gates, and a
repeat-bitwise stress of the recurrent kernel's fence-less exchange ring.  All through the C ABI."""
import multiprocessing as mp
import os

import numpy as np
import pytest

from flappie_b200 import signal as hs
from flappie_b200.api import Context, Model
from flappie_b200.model import KIND_GRU, KIND_LSTM, FlipflopModel, synthetic_reads

pytestmark = pytest.mark.gpu
TOL_TRANS = 1e-4            # north_star: intermediate floats within 1e-4

_ORC = {}


def _oracle_job(args):
    """(model name/seed/saturate, normalised signal, viterbi_only) -> dict; runs in a worker process"""
    key, sig, vit = args
    from oracle.pyoracle import Oracle
    if key not in _ORC:
        name, seed, sat = key
        kind, size, nbase = {"gru256": (KIND_GRU, 256, 4), "lstm384": (KIND_LSTM, 384, 4), "lstm256": (KIND_LSTM, 256, 4),
                             "gru256_5": (KIND_GRU, 256, 5)}[name]
        _ORC[key] = (Oracle(), FlipflopModel.synthetic(kind, size, nbase, seed, saturate=sat))
    orc, fm = _ORC[key]
    trans = orc.transitions(fm, sig, 1.0)
    out = {"trans": trans}
    score, path, qpath = orc.viterbi(trans)
    out["vit_path"] = path
    if not vit:
        out["fb"] = orc.basecall(fm, sig, 1.0, False)
    return out


def _pool_map(jobs):
    n = min(len(jobs), max(1, (os.cpu_count() or 2) - 1), 24)
    with mp.get_context("fork").Pool(n) as pool:
        return pool.map(_oracle_job, jobs, chunksize=1)


def _model(name, seed, sat=False):
    kind, size, nbase = {"gru256": (KIND_GRU, 256, 4), "lstm384": (KIND_LSTM, 384, 4), "lstm256": (KIND_LSTM, 256, 4),
                         "gru256_5": (KIND_GRU, 256, 5)}[name]
    return FlipflopModel.synthetic(kind, size, nbase, seed, saturate=sat)


@pytest.mark.parametrize("name,length", [("gru256", 50000), ("lstm384", 20000), ("lstm256", 30000)])
def test_longest_reads_against_the_oracle(gpu_lib, name, length):
    """cfg[3]'s longest read (50 k samples = 24 895 recurrent steps) and long LSTM reads: error growth of the fp16 hi/lo
    operands, the truncating accumulate and the MUFU gates over tens of thousands of steps stays inside 1e-4.  The
    --viterbi path is bit-exact on SHARED input (test_gpu_parity.py); decoded from the GPU's own trans -- up to 1e-4 away
    from the oracle's -- a near-tie can flip a short run of blocks, so the differing blocks are counted and bounded
    (profiles/r02_parity_report.txt: 0 to 10 of 24 896 on the 50 k read).  Run inside a ragged batch, as cfg[3] runs it."""
    fm = _model(name, 1)
    lens = [length, 1000, 2300, 7000, 12000, 3100, 40000 if length >= 40000 else 9000, 1700]
    raws = synthetic_reads(len(lens), lens, seed=5)
    sigs = [hs.prepare_read(r) for r in raws]
    m = Model(fm); ctx = Context(m)
    res = ctx.basecall(sigs, viterbi_only=True, want_trans=True)
    jobs = [((name, 1, False), sigs[i], True) for i in (0, 6)]
    for i, o in zip((0, 6), _pool_map(jobs)):
        d = float(np.max(np.abs(res.read_trans(i) - o["trans"])))
        assert d < TOL_TRANS, (name, lens[i], d)
        p, _ = res.read_path(i)
        nd = int(np.count_nonzero(p != o["vit_path"]))
        print(f"\n[parity] {name} {lens[i]} samples ({len(p) - 1} steps): max|d trans| {d:.2e}, differing Viterbi blocks {nd}/{len(p)}")
        assert nd <= len(p) // 1000, f"{name} {lens[i]} samples: {nd} of {len(p)} Viterbi blocks differ"
    ctx.close(); m.close()


@pytest.mark.parametrize("name,n,nsamp", [("gru256", 256, 1200), ("lstm384", 256, 1500), ("gru256_5", 128, 1200)])
def test_measured_base_mismatch_rate(gpu_lib, name, n, nsamp):
    """Called bases against the oracle over hundreds of reads, both decoding modes, reported as MEASURED rates (printed; the
    4000-sample numbers are in profiles/r02_parity_report.txt).  The decoders are bit-exact on shared input; end to end the
    GPU's trans sit up to 1e-4 from the oracle's, which can move a near-tie: at most 1 read in 32 may differ anywhere in
    its --viterbi path, at most 1 in 64 in a called base of the default mode."""
    fm = _model(name, 2)
    raws = synthetic_reads(n, nsamp, seed=77)
    sigs = [hs.prepare_read(r) for r in raws]
    m = Model(fm); ctx = Context(m)
    res_v = ctx.basecall(sigs, viterbi_only=True, want_trans=True)
    res_f = ctx.basecall(sigs, viterbi_only=False)
    outs = _pool_map([((name, 2, False), s, False) for s in sigs])
    bad_v = bad_f = 0
    dmax = 0.0
    for i, o in enumerate(outs):
        dmax = max(dmax, float(np.max(np.abs(res_v.read_trans(i) - o["trans"]))))
        bad_v += int(not np.array_equal(res_v.read_path(i)[0], o["vit_path"]))
        bases, _ = gpu_lib.emit_bases(*res_f.read_path(i), fm.nbase)
        bad_f += int(bases != o["fb"]["basecall"])
    print(f"\n[parity] {name}: {n} reads x {nsamp} samples: max|d trans| {dmax:.2e}; reads with a differing Viterbi path "
          f"{bad_v}/{n}; reads with a differing base in default mode {bad_f}/{n}")
    assert dmax < TOL_TRANS
    assert bad_v <= n // 32
    assert bad_f <= n // 64
    ctx.close(); m.close()


def test_s512_tensor_path_against_the_fp32_kernels(gpu_lib):
    """r103_native's size (S = 512, 16-CTA clusters, lo weight plane in shared memory) has two independent implementations
    on the device: the tcgen05 recurrence + W-stationary / A-tile GEMMs, and the fp32 CUDA-core kernels (FFB_FLAG_FP32_SIMT:
    plain FMA chains in k order, exact expf -- the same code that is oracle-checked at every other size).  64 reads, both
    against each other and against the oracle on a sample of reads."""
    fm = FlipflopModel.synthetic(KIND_LSTM, 512, 4, seed=4)
    sigs = [hs.prepare_read(r) for r in synthetic_reads(64, 3000, seed=31)]
    m = Model(fm); ctx = Context(m)
    simt = ctx.basecall(sigs, viterbi_only=True, want_trans=True, fp32_simt=True)
    res = ctx.basecall(sigs, viterbi_only=True, want_trans=True)
    assert np.array_equal(simt.blk_off, res.blk_off)
    nb = int(res.blk_off[-1])
    d = float(np.max(np.abs(res.trans[:nb] - simt.trans[:nb])))
    same = sum(int(np.array_equal(res.read_path(i)[0], simt.read_path(i)[0])) for i in range(len(sigs)))
    from oracle.pyoracle import Oracle
    orc = Oracle()
    do = ds = 0.0
    for i in (0, 17, 63):
        t = orc.transitions(fm, sigs[i], 1.0)
        do = max(do, float(np.max(np.abs(res.read_trans(i) - t))))
        ds = max(ds, float(np.max(np.abs(simt.read_trans(i) - t))))
    print(f"\n[parity] lstm512: 64 reads x 3000 samples: max|d trans| tensor vs fp32 kernels {d:.2e}; identical Viterbi paths "
          f"{same}/64; vs oracle (3 reads): tensor {do:.2e}, fp32 kernels {ds:.2e}")
    assert d < TOL_TRANS and do < TOL_TRANS and ds < TOL_TRANS and same >= 62
    ctx.close(); m.close()


@pytest.mark.parametrize("name", ["gru256", "lstm256"])
def test_saturated_gates(gpu_lib, name):
    """A weight seed with gate biases in U(-4, 4) (trained gates saturate; the U(-0.1, 0.1) biases of the other tests keep
    every gate in its linear region): ex2.approx / rcp.approx at the ends of their range, states frozen by z ~ 1 or pinned
    near +-1, cell states that integrate for hundreds of steps."""
    fm = _model(name, 3, sat=True)
    sigs = [hs.prepare_read(r) for r in synthetic_reads(32, 2500, seed=13)]
    m = Model(fm); ctx = Context(m)
    simt = ctx.basecall(sigs, viterbi_only=True, want_trans=True, fp32_simt=True)        # the fp32 CUDA-core kernels, same model
    res = ctx.basecall(sigs, viterbi_only=True, want_trans=True, keep_layers=True)
    layers = [ctx.fetch_layer(1 + l) for l in range(5)]
    frac = float(np.mean([np.mean(np.abs(x) > 0.95) for x in layers]))
    slow = float(np.mean([np.mean(np.abs(np.diff(x[:1000], axis=0)) < 1e-3) for x in layers]))
    print(f"\n[parity] {name} saturated seed: {100 * frac:.1f} % of states beyond +-0.95, {100 * slow:.1f} % of state updates below 1e-3")
    assert frac > 0.01 or slow > 0.2, (frac, slow)       # the regime is really reached
    outs = _pool_map([((name, 3, True), s, True) for s in sigs])
    dmax, dsimt, ndmax = 0.0, 0.0, 0
    for i, o in enumerate(outs):
        dmax = max(dmax, float(np.max(np.abs(res.read_trans(i) - o["trans"]))))
        dsimt = max(dsimt, float(np.max(np.abs(simt.read_trans(i) - o["trans"]))))
        ndmax = max(ndmax, int(np.count_nonzero(res.read_path(i)[0] != o["vit_path"])))
    print(f"[parity] {name} saturated seed: max|d trans| tensor path {dmax:.2e}, fp32 CUDA-core path {dsimt:.2e}, "
          f"most differing Viterbi blocks in a read {ndmax}")
    # With forget gates pinned near 1 an LSTM cell integrates its input for hundreds of steps, and with it every rounding
    # error of the gate pre-activations: the fp32 CUDA-core kernels (plain FMA chains, exact expf) end ~1e-4 from the oracle
    # on this seed, the tensor path -- 22-bit operands, truncating accumulate: 2-3x their per-step error -- between 3e-4
    # and 1e-3 (profiles/r02_acc_comp.txt, where a compensation of the truncation bias was tried and not adopted).  The GRU
    # stays inside the north-star 1e-4.  Both decode the oracle's path.  DESIGN.md section 7 lists this as a known limit;
    # FFB_FLAG_FP32_SIMT is the way around it for such a model.
    bound = TOL_TRANS if name.startswith("gru") else 2e-3
    assert dmax < bound and ndmax <= 2, (dmax, dsimt, ndmax)
    ctx.close(); m.close()


def test_repeat_bitwise_stress_cfg1(gpu_lib):
    """50 runs of BASELINE configs[1] (1024 x 4000, GRU-256): every output bit identical every time.  The recurrent
    kernel's exchange ring is written and read by the async proxy without a proxy fence (rnn_tc.cu), the streamed input
    GEMM overwrites Xin in place behind the recurrence, and the groups race each other for the tensor pipe -- a lost
    ordering anywhere shows up as a flipped bit here."""
    fm = FlipflopModel.for_name("r941_native_gru", seed=1)
    raws = synthetic_reads(1024, 4000, seed=7)
    m = Model(fm)
    ctxs = [Context(m), Context(m)]
    ref = None
    for it in range(50):
        res = ctxs[it % 2].basecall_raw(raws, emit=True)
        nb = int(res.blk_off[-1]) + res.n_reads
        called = "\n".join("%s %s" % res.read_bases(i) for i in range(res.n_reads))      # bytes past a read's NUL are not output
        cur = (res.path[:nb].tobytes(), res.qpath[:nb].tobytes(), res.score.tobytes(), called)
        if ref is None:
            ref = cur
        assert cur == ref, f"run {it} differs"
    # the same with trans requested (adds the fp64 partition scan and the -logZ/T shift, which moves the posteriors' last bits:
    # its own reference)
    t0 = None
    for it in range(6):
        res = ctxs[it % 2].basecall_raw(raws, want_trans=True)
        nb = int(res.blk_off[-1]) + res.n_reads
        cur = (res.path[:nb].tobytes(), res.qpath[:nb].tobytes(), res.trans.tobytes())
        if t0 is None:
            t0 = cur
        assert cur == t0, f"run {it} (want_trans) differs"
    for c in ctxs:
        c.close()
    m.close()


@pytest.mark.parametrize("name,n,nsamp", [("gru256", 300, 3000), ("lstm384", 300, 3000), ("lstm256", 200, 2500), ("gru256_5", 2100, 1500)])
def test_streamed_schedule_is_bit_identical_to_the_sequential_one(gpu_lib, name, n, nsamp):
    """The input GEMM of layer l+1 streamed behind layer l's recurrence (programmatic dependent launch, per-tile progress
    counters, Xin overwritten IN PLACE) against the plain kernel-after-kernel schedule (FFB_FLAG_KEEP_LAYERS switches the
    streaming off): same arithmetic, so every bit of trans / path / qpath / score must agree -- for the 128-block tiles of
    K = 256, the 64-block tiles of K = 384 (LSTM-384), ragged lengths, and a batch of several rounds per slot."""
    fm = _model(name, 4)
    rng = np.random.default_rng(5)
    raws = [r[: int(rng.integers(nsamp // 3, nsamp + 1))] for r in synthetic_reads(n, nsamp, seed=29)]
    sigs = [s for s in (hs.prepare_read(r) for r in raws) if s is not None]
    m = Model(fm); ctx = Context(m)
    a = ctx.basecall(sigs, want_trans=True)                        # streamed
    b = ctx.basecall(sigs, want_trans=True, keep_layers=True)      # sequential
    c = ctx.basecall(sigs, want_trans=True)                        # streamed again (workspaces reused)
    nb = int(a.blk_off[-1]) + a.n_reads
    for x in (b, c):
        assert np.array_equal(a.trans, x.trans) and np.array_equal(a.path[:nb], x.path[:nb])
        assert a.qpath[:nb].tobytes() == x.qpath[:nb].tobytes() and a.score.tobytes() == x.score.tobytes()   # qpath[0] is NAN
    ctx.close(); m.close()
