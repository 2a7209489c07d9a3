"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs.  Integer outputs (Viterbi path, trace up to rounding) bit-exact on shared
input; floats within the tolerance stated next to each assert (north star: 1e-4)."""
import numpy as np
import pytest

from flappie_b200.api import Context, Model
from flappie_b200.model import KIND_GRU, KIND_LSTM, FlipflopModel, synthetic_reads
from ffb_testutil import norm_reads

pytestmark = pytest.mark.gpu

TOL_LAYER = 2e-5     # abs, recurrent layer outputs in (-1, 1)
TOL_TRANS = 1e-4     # abs, transition scores (5 * tanh, minus logZ/T): the north-star tolerance


def _rand_trans(rng, T, nr):
    return (rng.normal(size=(T, nr)) * 2.0).astype(np.float32)


def _network_like_trans(rng, T, nbase):
    """Scores as calculate_transitions produces them: 5 * tanh(.) minus logZ / T (float64 partition scan,
    reference layers.c:1035-1096), so that forward/backward sums stay O(10) as they do on real reads."""
    nstate, nr = 2 * nbase, 2 * nbase * (nbase + 1)
    raw = (5.0 * np.tanh(rng.normal(size=(T, nr)))).astype(np.float32)
    a = np.zeros(nstate)
    for t in range(T):
        c = raw[t].astype(np.float64)
        flip = c[:nbase * nstate].reshape(nbase, nstate)            # [dest b1][source]
        flop = c[nbase * nstate:]
        new = np.empty(nstate)
        new[:nbase] = np.logaddexp.reduce(flip + a[None, :], axis=1)
        new[nbase:] = np.logaddexp(a[nbase:] + flop[nbase:], a[:nbase] + flop[:nbase])
        a = new
    logz = np.logaddexp.reduce(a)
    return (raw - np.float32(logz / T)).astype(np.float32)


@pytest.mark.parametrize("nbase", [4, 5])
@pytest.mark.parametrize("T", [1, 2, 31, 32, 33, 500, 1895])
def test_viterbi_bit_exact(gpu_lib, oracle, nbase, T):
    rng = np.random.default_rng(100 * nbase + T)
    nr = 2 * nbase * (nbase + 1)
    trans = _rand_trans(rng, T, nr)
    s_o, p_o, q_o = oracle.viterbi(trans)
    s_g, p_g, q_g = gpu_lib.decode_crf_flipflop(trans)
    assert np.array_equal(p_g, p_o)
    assert s_g == s_o                                   # fp32 adds in the same order: bit-exact
    assert np.isnan(q_g[0]) and np.array_equal(q_g[1:], q_o[1:])


@pytest.mark.parametrize("nbase", [4, 5])
def test_viterbi_ties_first_max_wins(gpu_lib, oracle, nbase):
    # quantised scores force many exact ties: the reference's strict '>' visit order decides
    rng = np.random.default_rng(7)
    nr = 2 * nbase * (nbase + 1)
    trans = rng.integers(-2, 3, size=(700, nr)).astype(np.float32)
    s_o, p_o, q_o = oracle.viterbi(trans)
    s_g, p_g, q_g = gpu_lib.decode_crf_flipflop(trans)
    assert np.array_equal(p_g, p_o) and s_g == s_o
    zeros = np.zeros((50, nr), np.float32)
    s_o, p_o, _ = oracle.viterbi(zeros)
    s_g, p_g, _ = gpu_lib.decode_crf_flipflop(zeros)
    assert np.array_equal(p_g, p_o) and s_g == s_o


def test_viterbi_combine_stays(gpu_lib, oracle):
    rng = np.random.default_rng(3)
    trans = _rand_trans(rng, 300, 40)
    _, p_o, _ = oracle.viterbi(trans)
    _, p_g, _ = gpu_lib.decode_crf_flipflop(trans, combine_stays=True)
    assert np.array_equal(p_g, np.where(p_o < 4, p_o, -1))


@pytest.mark.parametrize("nbase", [4, 5])
@pytest.mark.parametrize("T", [1, 17, 400, 1895])
def test_transpost_and_trace(gpu_lib, oracle, nbase, T):
    rng = np.random.default_rng(nbase + T)
    nr = 2 * nbase * (nbase + 1)
    trans = _network_like_trans(rng, T, nbase)
    tp_o = oracle.transpost(trans, True)
    tp_g = gpu_lib.transpost_crf_flipflop(trans, True)
    # The GPU scans are shift-invariant (running-normalised, max-shifted sums); the reference folds
    # logsumexp sequentially in fp32.  On globally normalised scores both stay O(10) and agree to ~1e-5.
    tol = 2e-4
    assert np.max(np.abs(tp_g - tp_o)) < tol
    pr_g = gpu_lib.transpost_crf_flipflop(trans, False)
    assert np.max(np.abs(pr_g - np.exp(tp_o))) < 1.5 * tol   # = the log-space tolerance times p <= 1
    # trace on SHARED input (the oracle's probabilities): integers, allow the .5 rounding edge
    prob = np.exp(tp_o).astype(np.float32)
    tr_o = oracle.trace(prob)
    tr_g = gpu_lib.trace_from_posterior(prob)
    assert tr_g.shape == tr_o.shape
    assert np.max(np.abs(tr_g - tr_o)) <= 1
    assert np.mean(tr_g != tr_o) < 0.01


def test_exp_activation(gpu_lib):
    x = np.linspace(-20, 3, 4000, dtype=np.float32).reshape(100, 40)
    y = gpu_lib.exp_activation_inplace(x)
    assert np.allclose(y, np.exp(x), rtol=3e-7, atol=0)


MODELS = [
    ("gru64_5b", KIND_GRU, 64, 5),
    ("gru96_4b", KIND_GRU, 96, 4),
    ("lstm96_4b", KIND_LSTM, 96, 4),
    ("lstm128_4b", KIND_LSTM, 128, 4),
    ("gru256_4b", KIND_GRU, 256, 4),     # the shapes with the tcgen05 recurrent kernel
    ("lstm256_4b", KIND_LSTM, 256, 4),
    ("lstm384_4b", KIND_LSTM, 384, 4),   # 12-CTA clusters
    ("lstm512_4b", KIND_LSTM, 512, 4),   # r103_native: 16-CTA clusters, lo weight plane in shared memory; the fp32 path keeps
                                         # half of each weight slice in shared memory and streams the rest from L2
]


@pytest.mark.parametrize("fp32_simt", [False, True])
@pytest.mark.parametrize("name,kind,size,nbase", MODELS)
def test_network_layers_small(gpu_lib, oracle, name, kind, size, nbase, fp32_simt):
    fm = FlipflopModel.synthetic(kind, size, nbase, seed=11)
    # ragged batch incl. lengths hitting every stride residue and a read that is too short
    lens = [1790, 1791, 1792, 1793, 1794, 600, 90, 2990, 10]
    reads = []
    for i, n in enumerate(lens):
        reads.append(norm_reads(1, max(n, 600) + 210, seed=50 + i)[0][:n])
    m = Model(fm)
    ctx = Context(m)
    res = ctx.basecall(reads, viterbi_only=True, want_trans=True, keep_layers=True, fp32_simt=fp32_simt)
    conv_g = ctx.fetch_layer(0)
    layers_g = [ctx.fetch_layer(1 + l) for l in range(5)]
    for i, sig in enumerate(reads):
        o = oracle.transitions(fm, sig, 1.0, want_layers=True)
        b0, b1 = int(res.blk_off[i]), int(res.blk_off[i + 1])
        if o is None:
            assert b1 == b0 and np.isnan(res.score[i])
            continue
        trans_o, conv_o, layers_o = o
        assert b1 - b0 == trans_o.shape[0] == fm.nblock(len(sig))
        # CUDA-core convolutions accumulate like the oracle (1e-5); the tensor-core last convolution of the LSTM topology
        # (fp16 hi/lo operands, K = 320 in 20 truncating fp32 accumulations) stays within 2e-5 -- north-star tolerance 1e-4
        assert np.max(np.abs(conv_g[b0:b1] - conv_o)) < (1e-5 if (fp32_simt or kind == KIND_GRU) else 2e-5), f"conv read {i}"
        for l in range(5):
            d = np.max(np.abs(layers_g[l][b0:b1] - layers_o[l]))
            assert d < TOL_LAYER, f"layer {l} read {i}: {d}"
        d = np.max(np.abs(res.read_trans(i) - trans_o))
        assert d < TOL_TRANS, f"trans read {i}: {d}"
        # Viterbi on the GPU's own trans must equal the oracle's Viterbi on the same numbers
        s_o, p_o, q_o = oracle.viterbi(res.read_trans(i))
        p_g, q_g = res.read_path(i)
        assert np.array_equal(p_g, p_o) and res.score[i] == s_o
        assert np.array_equal(q_g[1:], q_o[1:])
    ctx.close(); m.close()


@pytest.mark.parametrize("name,kind,size,nbase", MODELS[:2] + MODELS[3:])  # incl. gru256 (tensor path)
def test_basecall_forward_backward_small(gpu_lib, oracle, name, kind, size, nbase):
    fm = FlipflopModel.synthetic(kind, size, nbase, seed=5)
    reads = norm_reads(6, 2000, seed=9)
    m = Model(fm); ctx = Context(m)
    res = ctx.basecall(reads, viterbi_only=False, want_trans=True, want_trace=True)
    nb_same = 0
    for i, sig in enumerate(reads):
        tp_o = oracle.transpost(res.read_trans(i), True)
        assert np.max(np.abs(res.read_tpost(i) - tp_o)) < 2e-4
        # decode of the GPU's own posteriors is bit-exact against the oracle on those numbers
        s_o, p_o, q_o = oracle.viterbi(res.read_tpost(i))
        p_g, q_g = res.read_path(i)
        assert np.array_equal(p_g, p_o) and res.score[i] == s_o
        tr_o = oracle.trace(np.exp(res.read_tpost(i)).astype(np.float32))
        assert np.max(np.abs(res.read_trace(i).astype(np.int32) - tr_o)) <= 1
        # end to end against the oracle run from the raw signal: bases must agree
        full = oracle.basecall(fm, sig, 1.0, False)
        bases_g, qual_g = gpu_lib.emit_bases(p_g, q_g, fm.nbase)
        nb_same += int(bases_g == full["basecall"])
    # base for base on every read of this set (measured over 64 reads x 4000 samples per model: 0 reads with a differing
    # base, profiles/r02_parity_report.txt; the rate over 256 reads is asserted in test_gpu_hardening.py)
    assert nb_same == len(reads), f"{len(reads) - nb_same} of {len(reads)} reads differ from the oracle in a called base"
    ctx.close(); m.close()


def test_calculate_transitions_dropin(gpu_lib, oracle):
    fm = FlipflopModel.synthetic(KIND_GRU, 96, 4, seed=2)
    m = Model(fm)
    m.register("r941_native")
    sig = norm_reads(1, 1500, seed=4)[0]
    t_g = gpu_lib.calculate_transitions(sig, 1.0, 0)
    t_o = oracle.transitions(fm, sig, 1.0)
    assert t_g.shape == t_o.shape
    assert np.max(np.abs(t_g - t_o)) < TOL_TRANS
    t_g2 = gpu_lib.calculate_transitions(sig, 0.7, 0)
    t_o2 = oracle.transitions(fm, sig, 0.7)
    assert np.max(np.abs(t_g2 - t_o2)) < 2 * TOL_TRANS
    assert gpu_lib.calculate_transitions(sig[:10], 1.0, 0) is None   # shorter than the filter window


def test_empty_batch(gpu_lib):
    fm = FlipflopModel.synthetic(KIND_GRU, 64, 4, seed=2)
    m = Model(fm); ctx = Context(m)
    res = ctx.basecall([], viterbi_only=True)
    assert res.n_reads == 0
    res = ctx.basecall([np.zeros(5, np.float32)], viterbi_only=True)
    assert res.nblock(0) == 0 and np.isnan(res.score[0])
    ctx.close(); m.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name,reverse", [("r941_native_gru", False), ("r941_5mC", True), ("r941_rna002", True)])
def test_device_emission_identical_to_host_emission(gpu_lib, name, reverse):
    """Bases and quality characters emitted on the device (emit.cu) against ffb_emit_bases on the downloaded
    path / qpath (reference src/decode.c:66-79, src/flappie.c:284-297, src/util.h:285-305) -- every read identical,
    >= 1000 reads over the three parametrisations, ragged lengths, --reverse and the 5-base alphabet included."""
    fm = FlipflopModel.for_name(name, seed=5)
    m = Model(fm); ctx = Context(m)
    rng = np.random.default_rng(17)
    raws = [r[: int(rng.integers(300, 1400))] for r in synthetic_reads(400, 1400, seed=23)]
    raws[7] = raws[7][:12]                                            # rejected read: no bases, empty strings
    for vit in (False, True):
        res = ctx.basecall_raw(raws, viterbi_only=vit, emit=True, reverse=reverse)
        nb_total = 0
        for i in range(len(raws)):
            T = res.nblock(i)
            if T <= 0:
                assert res.nbases[i] == 0
                continue
            p, q = res.read_path(i)
            want = gpu_lib.emit_bases(p, q, fm.nbase, reverse=reverse)
            assert res.read_bases(i) == want, (name, i)
            nb_total += len(want[0])
        assert nb_total > 10000
    ctx.close(); m.close()
