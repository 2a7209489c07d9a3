"""GPU parity of the on-device signal preparation (ffb_upload_raw: trimming by chunk MADs, med-MAD / delta
normalisation; reference src/flappie.c:251-259) against the host restatement (flappie_b200/signal.py, pinned to
the reference's own fixtures in tests/test_oracle.py), against the committed fixture derived from
src/test/*_signal.crp, and -- where oracle/_ref is present -- against the reference's object code."""
import numpy as np
import pytest

from flappie_b200.api import Context, Model
from flappie_b200.model import KIND_GRU, KIND_LSTM, FlipflopModel, synthetic_reads
from flappie_b200 import signal as hs

pytestmark = pytest.mark.gpu


def _raws(seed=3):
    rng = np.random.default_rng(seed)
    lens = [4000, 4001, 3999, 1234, 777, 12000, 250, 150, 99, 5, 50000]
    raws = synthetic_reads(len(lens), lens, seed=seed)
    # a stalled leader / open-pore tail (low variance) so that the MAD trimming actually moves the bounds
    raws[0][:700] = 200.0 + rng.normal(0, 0.05, 700).astype(np.float32)
    raws[5][-1500:] = 60.0 + rng.normal(0, 0.02, 1500).astype(np.float32)
    # ties: quantised signal
    raws[3] = np.round(raws[3]).astype(np.float32)
    return raws


@pytest.mark.parametrize("seg,trim", [((100, 0.0), (200, 10)), ((100, 0.3), (0, 0)), ((64, 0.05), (30, 500))])
def test_device_trim_and_normalise_match_host(gpu_lib, seg, trim):
    fm = FlipflopModel.synthetic(KIND_GRU, 64, 4, seed=2)
    m = Model(fm); ctx = Context(m)
    raws = _raws()
    res = ctx.basecall_raw(raws, viterbi_only=True, trim=trim, segmentation=seg)
    kept = []
    for i, r in enumerate(raws):
        se = hs.trim_and_segment_raw(r, trim[0], trim[1], seg[0], seg[1])
        s, e = int(res.start[i]), int(res.end[i])
        if se is None:
            assert s >= e, f"read {i}: host drops it, device kept [{s}, {e})"
            assert res.nblock(i) == 0 and np.isnan(res.score[i])
            continue
        assert (s, e) == se, f"read {i}"                     # integers: exact
        kept.append(hs.medmad_normalise_array(r[s:e]))
        assert res.nblock(i) == max(fm.nblock(e - s), 0)
    want = np.concatenate(kept)
    got = ctx.fetch_signal(want.shape[0])
    assert np.array_equal(got, want), f"max diff {np.max(np.abs(got - want))}"   # same IEEE arithmetic: bit-exact
    ctx.close(); m.close()


def test_device_delta_path(gpu_lib):
    """--delta: difference_array then / delta, no med-MAD (reference src/flappie.c:257-258)."""
    fm = FlipflopModel.synthetic(KIND_LSTM, 96, 4, seed=2)
    m = Model(fm); ctx = Context(m)
    raws = synthetic_reads(5, [4000, 3000, 2345, 800, 10000], seed=8)
    res = ctx.basecall_raw(raws, viterbi_only=True, delta=1.5)
    kept = []
    for i, r in enumerate(raws):
        s, e = int(res.start[i]), int(res.end[i])
        assert (s, e) == hs.trim_and_segment_raw(r)
        kept.append((hs.difference_array(r[s:e]) / np.float32(1.5)).astype(np.float32))
    want = np.concatenate(kept)
    assert np.array_equal(ctx.fetch_signal(want.shape[0]), want)
    ctx.close(); m.close()


def test_raw_path_equals_host_prepared_path(gpu_lib):
    """basecall_raw(raw) == basecall(prepare_read(raw)): same bits in, same kernels after."""
    fm = FlipflopModel.synthetic(KIND_GRU, 96, 4, seed=5)
    m = Model(fm); ctx = Context(m)
    raws = synthetic_reads(6, [4000, 4000, 2500, 7000, 1500, 4000], seed=12)
    a = ctx.basecall_raw(raws, want_trans=True)
    b = ctx.basecall([hs.prepare_read(r) for r in raws], want_trans=True)
    assert np.array_equal(a.blk_off, b.blk_off)
    nb = int(a.blk_off[-1])
    assert np.array_equal(a.trans[:nb], b.trans[:nb])
    assert np.array_equal(a.path[:nb + 6], b.path[:nb + 6]) and np.array_equal(a.score, b.score)
    ctx.close(); m.close()


def test_reference_fixture_on_device(gpu_lib):
    """tests/golden/signal_fixture.npz: head of the reference's raw_signal.crp, its trim bounds and the
    normalised values of normalised_signal.crp (tolerance of src/test/test_flappie_signal.c:109)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "signal_fixture.npz"))
    fm = FlipflopModel.synthetic(KIND_GRU, 64, 4, seed=2)
    m = Model(fm); ctx = Context(m)
    res = ctx.basecall_raw([g["raw_pa_head"]], viterbi_only=True)
    assert (int(res.start[0]), int(res.end[0])) == (int(g["start"]), int(g["end"]))
    got = ctx.fetch_signal(int(g["end"]) - int(g["start"]))
    assert np.max(np.abs(got - g["normalised_head"])) < 1e-5
    ctx.close(); m.close()


def test_device_matches_reference_object_code(gpu_lib, ref):
    """Where oracle/_ref is present: the reference's own medmad_normalise_array on the device's kept range."""
    fm = FlipflopModel.synthetic(KIND_GRU, 64, 4, seed=2)
    m = Model(fm); ctx = Context(m)
    raws = synthetic_reads(3, [4000, 9000, 2000], seed=21)
    res = ctx.basecall_raw(raws, viterbi_only=True)
    want = np.concatenate([ref.medmad_normalise(r[int(res.start[i]):int(res.end[i])]) for i, r in enumerate(raws)])
    got = ctx.fetch_signal(want.shape[0])
    assert np.max(np.abs(got - want)) < 1e-6
    ctx.close(); m.close()


def test_begin_finish_and_reserve_do_not_change_results(gpu_lib):
    """ffb_reserve (workspaces sized up front) and the ffb_submit_raw_begin / ffb_submit_raw_finish split (what the
    `flappie --devices` pipeline calls) give the bytes of the one-call path."""
    from flappie_b200.api import FLAG_WANT_TRANS
    fm = FlipflopModel.synthetic(KIND_GRU, 96, 4, seed=5)
    m = Model(fm)
    raws = synthetic_reads(40, [4000, 2500, 7000, 1500, 300, 4000, 9000, 123] * 5, seed=21)
    ctx = Context(m)
    want = ctx.basecall_raw(raws, want_trans=True, emit=True)
    ctx.close()

    ctx = Context(m)
    ctx.reserve(64, 8192, FLAG_WANT_TRANS)
    assert ctx.total_blocks() == 0                                # a reservation leaves no batch behind
    lens = np.array([len(r) for r in raws], np.int64)
    raw_off = np.zeros(len(raws) + 1, np.int64)
    np.cumsum(lens, out=raw_off[1:])
    raw = np.concatenate(raws).astype(np.float32)
    for _ in range(2):                                            # twice: the second batch reuses every buffer
        b, o = ctx.make_batch(raw, raw_off, 1.0, FLAG_WANT_TRANS, emit=True)
        rb, start, end = ctx.make_raw_batch(raw, raw_off)
        ctx.submit_raw_begin(rb, b)
        ctx.submit_raw_finish(b)
        ctx.collect(b)
        assert np.array_equal(start[:len(raws)], want.start) and np.array_equal(end[:len(raws)], want.end)
        assert np.array_equal(o["blk_off"], want.blk_off)
        nb = int(want.blk_off[-1])
        assert o["trans"][:nb].tobytes() == want.trans[:nb].tobytes()
        assert np.array_equal(o["path"][:nb + len(raws)], want.path[:nb + len(raws)])
        assert o["score"].tobytes() == want.score.tobytes()
        assert np.array_equal(o["nbases"], want.nbases)
        for i in range(len(raws)):
            s, k = int(want.blk_off[i]) + i, int(want.nbases[i])
            assert o["bases"][s:s + k].tobytes() == want.bases[s:s + k].tobytes()
            assert o["quals"][s:s + k].tobytes() == want.quals[s:s + k].tobytes()
    # finish without begin is an error, not a crash
    b, o = ctx.make_batch(raw, raw_off, 1.0, 0)
    with pytest.raises(Exception):
        ctx.submit_raw_finish(b)
    ctx.close(); m.close()
