"""CPU tests that PIN the oracle (oracle/flappie_oracle.c):
  * against golden vectors generated from the reference's own object code
    (tests/golden/*.npz, tests/golden/make_golden.py) -- always;
  * against that object code directly (oracle/_ref) -- wherever it has been built;
  * against the reference test-suite's own fixtures (src/test/*.crp) -- where
    /root/reference is mounted.
"""
import os

import numpy as np
import pytest

from flappie_b200.model import KIND_GRU, KIND_LSTM, FlipflopModel
from oracle.pyoracle import ACT_NONE, ACT_SWISH, ACT_TANH

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


# ---- golden vectors ------------------------------------------------------------------
def test_conv_golden(oracle):
    g = gold("conv_golden.npz")
    keys = sorted(k[:-2] for k in g.files if k.endswith("_x"))
    assert len(keys) >= 9
    for k in keys:
        stride = int(k.split("_")[0][1:])
        y = oracle.convolution(g[k + "_x"], g[k + "_W"], g[k + "_b"], stride, ACT_NONE)
        assert y.shape == g[k + "_y"].shape, k
        assert np.max(np.abs(y - g[k + "_y"])) < 3e-5, k   # SGEMM summation order only


def test_conv_right_edge_quirk_is_reproduced(oracle):
    # reference layers.c:229-271: with T % stride == 0 the last columns are NOT the textbook
    # "same" convolution (SURVEY.md section 0.4 worked example): last column is bias only.
    g = gold("conv_golden.npz")
    k = "s5_w19_nf16_f8_T100"
    y = g[k + "_y"]
    assert np.allclose(y[-1], g[k + "_b"], atol=0)            # bias-only final column
    k = "s2_w19_nf1_f8_T100"
    assert np.allclose(g[k + "_y"][-1], g[k + "_b"], atol=0)
    k = "s2_w19_nf1_f8_T101"                                   # odd T: textbook
    assert not np.allclose(g[k + "_y"][-1], g[k + "_b"], atol=1e-6)


def test_conv_plan_matches_textbook_away_from_the_tail(oracle):
    for T in (100, 101, 3790, 3793):
        for stride, winlen in ((2, 19), (5, 19), (1, 5)):
            terms, ncol = oracle.conv_plan(T, winlen, stride)
            padL = (winlen - 1) // 2
            per_col = {}
            for c, xs, tl, nt in terms:
                per_col.setdefault(c, []).append((xs, tl, nt))
            for c in range(ncol - 32):
                xs, tl, nt = c * stride - padL, 0, winlen
                if xs < 0:
                    tl, nt, xs = -xs, winlen + xs, 0
                assert per_col[c] == [(xs, tl, nt)], (T, stride, c)
    assert oracle.conv_plan(10, 19, 2)[0] is None              # shorter than the window: rejected


def test_decode_golden(oracle):
    g = gold("decode_golden.npz")
    for nb in (4, 5):
        k = f"b{nb}"
        score, path, qpath = oracle.viterbi(g[k + "_trans"])
        assert np.array_equal(path, g[k + "_path"]) and score == g[k + "_score"]
        assert np.array_equal(qpath[1:], g[k + "_qpath"][1:]) and np.isnan(qpath[0])
        tpost = oracle.transpost(g[k + "_trans"], True)
        assert np.array_equal(tpost, g[k + "_tpost"])           # same libm, same fold order: bit-exact
        s2, p2, _ = oracle.viterbi(tpost)
        assert np.array_equal(p2, g[k + "_post_path"]) and s2 == g[k + "_post_score"]
        assert np.array_equal(oracle.trace(np.exp(tpost).astype(np.float32)), g[k + "_trace"])


@pytest.mark.parametrize("name", ["gru64_5b", "gru96_4b", "lstm96_4b"])
def test_network_golden(oracle, name):
    g = gold(f"net_{name}.npz")
    fm = FlipflopModel.synthetic(int(g["kind"]), int(g["size"]), int(g["nbase"]), seed=int(g["seed"]))
    trans, conv, layers = oracle.transitions(fm, g["signal"], 1.0, want_layers=True)
    gc = g["conv"].astype(np.float32)                                     # stored as fp16: 2^-11 relative (swish is unbounded)
    assert np.max(np.abs(conv - gc) / np.maximum(1.0, np.abs(gc))) < 2e-3
    assert np.max(np.abs(layers[0] - g["layer1"].astype(np.float32))) < 2e-3
    assert np.max(np.abs(layers[4] - g["layer5"])) < 2e-5
    assert np.max(np.abs(trans - g["trans"])) < 1e-4
    # decoding the GOLDEN trans must reproduce the golden calls exactly
    score, path, qpath = oracle.viterbi(g["trans"])
    assert np.array_equal(path, g["vit_path"]) and score == g["vit_score"]
    bases, qual = oracle.emit_bases(path, qpath, fm.nbase)
    assert bases == str(g["vit_bases"]) and qual == str(g["vit_qual"])
    # and the oracle run end to end from the signal calls the same bases
    for vo, key in ((True, "vit"), (False, "fb")):
        out = oracle.basecall(fm, g["signal"], 1.0, vo, want_trace=True)
        assert out["basecall"] == str(g[key + "_bases"])
        assert out["quality"] == str(g[key + "_qual"])
        assert np.array_equal(out["path"], g[key + "_path"])
    out = oracle.basecall(fm, g["signal"], 1.0, False, want_trace=True)
    assert np.max(np.abs(out["trace"].astype(np.int32) - g["fb_trace"].astype(np.int32))) <= 1


def test_signal_fixture_host_prep():
    from flappie_b200.signal import medmad_normalise_array, trim_and_segment_raw
    g = gold("signal_fixture.npz")
    se = trim_and_segment_raw(g["raw_pa_head"])
    assert se == (int(g["start"]), int(g["end"]))
    x = medmad_normalise_array(g["raw_pa_head"][se[0]:se[1]])
    assert np.max(np.abs(x - g["normalised_head"])) < 1e-5       # tolerance of test_flappie_signal.c:109


# ---- against the reference's object code (where built) ---------------------------------
def test_oracle_vs_ref_pieces(oracle, ref):
    rng = np.random.default_rng(0)
    S, T = 64, 211
    X = rng.uniform(-1, 1, (T, S)).astype(np.float32)
    for G in (3, 4):
        iW = rng.uniform(-.2, .2, (G * S, S)).astype(np.float32)
        sW = rng.uniform(-.2, .2, (G * S, S)).astype(np.float32)
        b = rng.uniform(-.3, .3, G * S).astype(np.float32)
        xin = ref.affine(X, iW, b)
        assert np.max(np.abs(oracle.affine(X, iW, b) - xin)) < 1e-5
        for bw in (0, 1):
            if G == 3:
                assert np.max(np.abs(oracle.grumod(xin, sW, bw) - ref.grumod(xin, sW, bw))) < 2e-6
            else:
                assert np.max(np.abs(oracle.lstm(xin, sW, bw) - ref.lstm(xin, sW, bw))) < 2e-6
    for nbase in (4, 5):
        nr = 2 * nbase * (nbase + 1)
        W = rng.uniform(-.3, .3, (nr, S)).astype(np.float32)
        b = rng.uniform(-.3, .3, nr).astype(np.float32)
        for temp in (1.0, 0.6):
            ta, _ = oracle.globalnorm(X, W, b, temp)
            tr = ref.globalnorm(X, W, b, temp)
            assert np.max(np.abs(ta - tr)) < 1e-5
        so, po, qo = oracle.viterbi(tr)
        sr, pr, qr = ref.viterbi(tr)
        assert so == sr and np.array_equal(po, pr) and np.array_equal(qo[1:], qr[1:])
        assert np.array_equal(oracle.transpost(tr, True), ref.transpost(tr, True))
        prob = np.exp(ref.transpost(tr, True)).astype(np.float32)
        assert np.array_equal(oracle.trace(prob), ref.trace(prob))


def test_oracle_vs_ref_conv_sweep(oracle, ref):
    rng = np.random.default_rng(1)
    for stride in (1, 2, 3, 5):
        for nf, nfilter in ((1, 4), (4, 16), (16, 4)):
            for T in list(range(19, 64)) + [100, 101, 3790, 3791, 3795]:
                x = rng.normal(size=(T, nf)).astype(np.float32)
                W = (rng.normal(size=(nfilter, 19, nf)) * .3).astype(np.float32)
                b = rng.normal(size=(nfilter,)).astype(np.float32)
                a = oracle.convolution(x, W, b, stride, ACT_TANH)
                assert a is not None
                assert np.max(np.abs(a - ref.convolution(x, W, b, stride, ACT_TANH))) < 3e-5, (stride, nf, T)


@pytest.mark.parametrize("kind,size,nbase", [(KIND_GRU, 64, 4), (KIND_LSTM, 96, 4)])
def test_oracle_vs_ref_network(oracle, ref, kind, size, nbase):
    from ffb_testutil import norm_reads
    fm = FlipflopModel.synthetic(kind, size, nbase, seed=3)
    sig = norm_reads(1, 1600, seed=8)[0]
    rm = ref.model(fm)
    ta, ca, la = oracle.transitions(fm, sig, 1.0, True)
    tr, cr, lr = ref.transitions(rm, sig, 1.0, True)
    assert np.max(np.abs(ca - cr)) < 1e-5
    for a, b in zip(la, lr):
        assert np.max(np.abs(a - b)) < 5e-6
    assert np.max(np.abs(ta - tr)) < 5e-5
    for vo in (True, False):
        oa = oracle.basecall(fm, sig, 1.0, vo)
        ob = ref.basecall(rm, sig, 1.0, vo)
        assert oa["basecall"] == ob["basecall"] and oa["quality"] == ob["quality"]


def test_reference_signal_fixtures(ref):
    """The reference test-suite's own vectors (src/test/test_flappie_signal.c:67-111)."""
    tdir = "/root/reference/src/test"
    if not os.path.isdir(tdir):
        pytest.skip("reference test fixtures not mounted")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    from flappie_b200.signal import medmad_normalise_array, trim_raw_by_mad
    raw = mg.read_crp(os.path.join(tdir, "raw_signal.crp"))[:, 0]
    trimmed = mg.read_crp(os.path.join(tdir, "trimmed_signal.crp"))[:, 0]
    normalised = mg.read_crp(os.path.join(tdir, "normalised_signal.crp"))[:, 0]
    unit = np.float32(1373.41) / np.float32(8192.0)
    pa = ((raw + np.float32(16.0)) * unit).astype(np.float32)
    start, end = trim_raw_by_mad(pa, 0, pa.shape[0], 100, 0.0)
    assert start == 0 and end == (pa.shape[0] // 100) * 100
    mine = pa[start + 200:end - 10]
    assert mine.shape == trimmed.shape and np.max(np.abs(mine - trimmed)) < 1e-4
    assert np.max(np.abs(medmad_normalise_array(trimmed) - normalised)) < 1e-5
    assert np.max(np.abs(ref.medmad_normalise(trimmed) - normalised)) < 1e-5


def test_host_signal_prep_vs_reference_object_code(ref):
    """flappie_b200/signal.py (the host restatement the device kernels are tested against) versus the reference's
    own quantilef / madf / medmad_normalise_array (src/util.c:100-212): bit-exact."""
    from flappie_b200 import signal as hs
    rng = np.random.default_rng(17)
    for n in (2, 3, 40, 41, 100, 379, 3790):
        x = rng.normal(90, 12, n).astype(np.float32)
        if n == 100:
            x = np.round(x)                      # ties
        for p in (0.0, 0.05, 0.3, 0.5, 0.77, 1.0):
            assert hs.quantilef(x, p) == ref.quantile(x, p), (n, p)
        assert hs.madf(x) == ref.mad(x), n
        assert np.array_equal(hs.medmad_normalise_array(x), ref.medmad_normalise(x)), n


def test_runlength_head_oracle_vs_reference_object_code(oracle, ref):
    """The run-length ("runnie") head restated in oracle/flappie_oracle.c against the reference's own
    globalnorm_runlengthV2 / decode_crf_runlength / transpost_crf_runlength (layers.c:1255-1358, decode.c:901-1159)."""
    rng = np.random.default_rng(23)
    for T in (1, 2, 57, 400):
        S, nr = 32, 40
        h = rng.uniform(-1, 1, (T, S)).astype(np.float32)
        W = (rng.normal(size=(nr, S)) * 0.4).astype(np.float32)
        b = (rng.normal(size=nr) * 0.2).astype(np.float32)
        for temperature in (1.0, 0.7):
            p_o, _ = oracle.globalnorm_runlength(h, W, b, temperature)
            p_r = ref.globalnorm_runlength(h, W, b, temperature)
            assert p_o.shape == p_r.shape == (T, nr)
            assert np.max(np.abs(p_o - p_r)) < 2e-5          # OpenBLAS vs increasing-k summation in the affine map
        s_o, path_o = oracle.rle_viterbi(p_r)                 # shared input: integers and the score bit-exact
        s_r, path_r = ref.rle_viterbi(p_r)
        assert np.array_equal(path_o, path_r) and s_o == s_r
        post_o, post_r = oracle.rle_transpost(p_r), ref.rle_transpost(p_r)
        assert np.array_equal(post_o, post_r)                 # same libm, same fold order
        s_o, path_o = oracle.rle_viterbi(post_r)
        s_r, path_r = ref.rle_viterbi(post_r)
        assert np.array_equal(path_o, path_r) and s_o == s_r
        bases, shape, scale, dwell = oracle.emit_runs(path_o, post_r)
        assert len(bases) == int(np.sum(path_o < 4)) and int(dwell.sum()) <= T
        assert np.all(shape >= 1.0) and np.all(scale > 0.0)
    # quantised scores: ties resolved in the reference's visit order
    q = rng.integers(-2, 3, size=(300, 40)).astype(np.float32)
    assert np.array_equal(oracle.rle_viterbi(q)[1], ref.rle_viterbi(q)[1])


def test_runlength_and_signal_golden(oracle):
    """tests/golden/rle_golden.npz (generated from the reference's object code by make_golden_rle.py): pins the
    run-length head of the oracle and the host signal prep where oracle/_ref itself is absent."""
    from flappie_b200 import signal as hs
    g = gold("rle_golden.npz")
    for temp in (1.0, 0.7):
        p, _ = oracle.globalnorm_runlength(g["h"], g["W"], g["b"], temp)
        assert np.max(np.abs(p - g[f"param_t{temp}"])) < 2e-5
    p = g["param_t1.0"]
    s, path = oracle.rle_viterbi(p)
    assert np.array_equal(path, g["vit_path"]) and s == g["vit_score"]
    post = oracle.rle_transpost(p)
    assert np.array_equal(post, g["post"])
    s2, path2 = oracle.rle_viterbi(post)
    assert np.array_equal(path2, g["post_path"]) and s2 == g["post_score"]
    assert np.array_equal(oracle.rle_viterbi(g["tie_param"])[1], g["tie_path"])
    x = g["sig"]
    for p_, want in zip((0.0, 0.05, 0.3, 0.5, 0.77, 1.0), g["sig_q"]):
        assert hs.quantilef(x, p_) == want
    assert hs.madf(x) == g["sig_mad"] and np.array_equal(hs.medmad_normalise_array(x), g["sig_norm"])
