"""GPU parity of the run-length ("runnie") head: globalnorm_runlengthV2, decode_crf_runlength,
transpost_crf_runlength and runnie's run loop against the plain-C oracle (pinned to the reference's object code in
tests/test_oracle.py::test_runlength_head_oracle_vs_reference_object_code)."""
import numpy as np
import pytest

from flappie_b200.api import Context, Model
from flappie_b200.model import KIND_LSTM, FlipflopModel
from ffb_testutil import norm_reads

pytestmark = pytest.mark.gpu


def _params(rng, T):
    p = rng.normal(size=(T, 40)).astype(np.float32) * 2.0
    p[:, :8] = 1.0 + np.abs(p[:, :8])
    return p


@pytest.mark.parametrize("T", [1, 2, 33, 500, 1895])
def test_rle_viterbi_bit_exact(gpu_lib, oracle, T):
    p = _params(np.random.default_rng(T), T)
    s_o, path_o = oracle.rle_viterbi(p)
    s_g, path_g = gpu_lib.decode_crf_runlength(p)
    assert np.array_equal(path_g, path_o) and s_g == s_o


def test_rle_viterbi_ties(gpu_lib, oracle):
    q = np.random.default_rng(7).integers(-2, 3, size=(700, 40)).astype(np.float32)
    s_o, path_o = oracle.rle_viterbi(q)
    s_g, path_g = gpu_lib.decode_crf_runlength(q)
    assert np.array_equal(path_g, path_o) and s_g == s_o
    z = np.zeros((40, 40), np.float32)
    assert np.array_equal(gpu_lib.decode_crf_runlength(z)[1], oracle.rle_viterbi(z)[1])


@pytest.mark.parametrize("T", [1, 17, 400, 1895])
def test_rle_transpost(gpu_lib, oracle, T):
    rng = np.random.default_rng(100 + T)
    p = _params(rng, T)
    p[:, 8:] = (5.0 * np.tanh(rng.normal(size=(T, 32)))).astype(np.float32) - np.float32(4.6)   # ~ globally normalised
    post_o = oracle.rle_transpost(p)
    post_g = gpu_lib.transpost_crf_runlength(p)
    assert np.array_equal(post_g[:, :8], p[:, :8])                         # shape / scale rows copied through
    assert np.max(np.abs(post_g - post_o)) < 1e-4 * max(1.0, np.max(np.abs(post_o)) / 10)
    # decoding the GPU's own posteriors agrees with the oracle's decoder on those numbers
    assert np.array_equal(gpu_lib.decode_crf_runlength(post_g)[1], oracle.rle_viterbi(post_g)[1])


@pytest.mark.parametrize("size,fp32_simt", [(96, False), (256, False), (256, True)])
def test_runlength_network_end_to_end(gpu_lib, oracle, size, fp32_simt):
    fm = FlipflopModel.synthetic(KIND_LSTM, size, 4, seed=7)
    fm.head = "runlength"
    reads = norm_reads(5, 2500, seed=19) + [norm_reads(1, 900, seed=3)[0][:400]]
    m = Model(fm); ctx = Context(m)
    a = ctx.basecall(reads, viterbi_only=True, want_trans=True, fp32_simt=fp32_simt)
    b = ctx.basecall(reads, viterbi_only=False, want_trans=True, fp32_simt=fp32_simt)
    for i, sig in enumerate(reads):
        p_o = oracle.runlength_transitions(fm, sig, 1.0)
        assert p_o.shape == a.read_trans(i).shape
        assert np.max(np.abs(a.read_trans(i) - p_o)) < 1e-4, f"read {i}"
        # --viterbi: decode of the GPU's own parameters, bit-exact against the oracle's decoder
        s_o, path_o = oracle.rle_viterbi(a.read_trans(i))
        st, rle = a.read_rle(i)
        assert np.array_equal(st, path_o) and a.score[i] == s_o
        assert np.array_equal(rle, a.read_trans(i)[:, :8])
        # default mode: posteriors, then Viterbi on them
        post_o = oracle.rle_transpost(b.read_trans(i))
        assert np.max(np.abs(b.read_tpost(i) - post_o)) < 2e-4
        st_b, rle_b = b.read_rle(i)
        assert np.array_equal(st_b, oracle.rle_viterbi(b.read_tpost(i))[1])
        # runs: same text as the oracle's loop over the same path / parameters
        bases_g, shape_g, scale_g, dwell_g = gpu_lib.emit_runs(st_b, rle_b)
        bases_o, shape_o, scale_o, dwell_o = oracle.emit_runs(st_b, b.read_tpost(i))
        assert bases_g == bases_o and np.array_equal(dwell_g, dwell_o)
        assert np.array_equal(shape_g, shape_o) and np.array_equal(scale_g, scale_o)
    # default mode without want_trans (no logZ pass): same states, shapes, scales
    c = ctx.basecall(reads, fp32_simt=fp32_simt)
    for i in range(len(reads)):
        assert np.array_equal(c.read_rle(i)[0], b.read_rle(i)[0]) and np.array_equal(c.read_rle(i)[1], b.read_rle(i)[1])
    ctx.close(); m.close()
