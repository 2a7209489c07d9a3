"""BASELINE.json configs at their own sizes, through the C ABI.  The oracle cannot finish thousands of S=256
reads in seconds, so each config is checked by (a) the oracle / reference object code on a few of its reads and
(b) size-independent properties of the path: reads are independent end to end, hence every per-read result must
be BIT-IDENTICAL whatever batch the read travels in (batch-composition, permutation and shard invariance), and a
second pass over the same batch must reproduce the first (idempotence / no stale state in the workspaces)."""
import numpy as np
import pytest

from flappie_b200 import signal as hs
from flappie_b200.api import Context, Model
from flappie_b200.model import FlipflopModel, synthetic_reads
from flappie_b200.shard import shard_reads

pytestmark = pytest.mark.gpu
TOL_TRANS = 1e-4     # north-star tolerance on intermediate floats


def _per_read(res, i):
    p, q = res.read_path(i)
    return p.copy(), q.copy(), np.float32(res.score[i])


def _same(a, b):
    return np.array_equal(a[0], b[0]) and np.array_equal(a[1][1:], b[1][1:]) and a[2] == b[2]


def test_cfg1_1024_reads_r941_native_idempotent_and_permutation_invariant(gpu_lib, oracle):
    """configs[1]: 1024 synthetic 4000-sample reads, r941_native (north-star GRU-256 topology)."""
    fm = FlipflopModel.for_name("r941_native_gru", seed=1)
    raws = synthetic_reads(1024, 4000, seed=7)
    reads = [hs.prepare_read(r) for r in raws]
    m = Model(fm); ctx = Context(m)
    a = ctx.basecall(reads)
    b = ctx.basecall(reads)                                    # second pass over warm workspaces
    assert np.array_equal(a.path, b.path) and np.array_equal(a.score, b.score)
    assert np.array_equal(a.qpath[~np.isnan(a.qpath)], b.qpath[~np.isnan(b.qpath)])
    perm = np.random.default_rng(0).permutation(1024)
    c = ctx.basecall([reads[j] for j in perm])
    for k in range(0, 1024, 37):
        assert _same(_per_read(a, int(perm[k])), _per_read(c, k)), f"read {perm[k]} changed with its position in the batch"
    # a small batch of the same reads (different cluster / group / tile assignment) gives the same bits
    sub = [3, 500, 1023]
    d = ctx.basecall([reads[j] for j in sub])
    for k, j in enumerate(sub):
        assert _same(_per_read(a, j), _per_read(d, k))
    # (want_trans adds the -logZ/T shift of globalnorm to `trans`: mathematically neutral for the posteriors,
    # not bitwise -- so the oracle comparison below is a separate call)
    d = ctx.basecall([reads[j] for j in sub], want_trans=True)
    # and the oracle agrees on one full-size read: floats within the north-star tolerance, bases identical
    t_o = oracle.transitions(fm, reads[3], 1.0)
    assert np.max(np.abs(d.read_trans(0) - t_o)) < TOL_TRANS
    full = oracle.basecall(fm, reads[3], 1.0, False)
    bases, _ = gpu_lib.emit_bases(*d.read_path(0), fm.nbase)
    assert bases == full["basecall"]
    ctx.close(); m.close()


def test_cfg2_4096_reads_r941_5mC(gpu_lib, oracle):
    """configs[2]: --model r941_5mC (5-base CpG head, 60 transition rows), 4096 reads on one GPU."""
    fm = FlipflopModel.for_name("r941_5mC", seed=1)
    assert fm.nparam == 60 and fm.nbase == 5
    base = synthetic_reads(256, 4000, seed=11)
    prepared = [hs.prepare_read(r) for r in base]
    reads = [prepared[i % 256] for i in range(4096)]            # 16 copies of 256 distinct reads
    m = Model(fm); ctx = Context(m)
    a = ctx.basecall(reads)
    assert a.n_reads == 4096 and int(a.blk_off[-1]) == sum(fm.nblock(len(r)) for r in reads)
    # every copy of a read, in whatever wave / cluster / group it landed, decodes to the same bits
    for i in range(0, 256, 17):
        ref_i = _per_read(a, i)
        for rep in range(1, 16):
            assert _same(ref_i, _per_read(a, i + 256 * rep)), f"read {i} copy {rep}"
    d = ctx.basecall([reads[5]])
    assert _same(_per_read(a, 5), _per_read(d, 0))
    d = ctx.basecall([reads[5]], want_trans=True)
    t_o = oracle.transitions(fm, reads[5], 1.0)
    assert np.max(np.abs(d.read_trans(0) - t_o)) < TOL_TRANS
    bases, _ = gpu_lib.emit_bases(*d.read_path(0), fm.nbase)
    assert set(bases) <= set("ACGTZ")
    assert bases == oracle.basecall(fm, reads[5], 1.0, False)["basecall"]
    ctx.close(); m.close()


def test_cfg3_r10C_pcr_mixed_lengths_sharded(gpu_lib, oracle):
    """configs[3]: --model r10C_pcr, lengths log-uniform 1k-50k, read-sharded over 8 ranks (emulated one rank at
    a time on this GPU: the shards share nothing, so rank r's output must equal the unsharded batch's)."""
    fm = FlipflopModel.for_name("r10C_pcr", seed=1)
    rng = np.random.default_rng(5)
    lens = np.exp(rng.uniform(np.log(1000), np.log(50000), size=96)).astype(np.int64)
    raws = synthetic_reads(96, lens, seed=13)
    m = Model(fm); ctx = Context(m)
    whole = ctx.basecall_raw(raws)                              # raw path: trimming + normalisation on the device
    shards = shard_reads(lens, 8)
    assert sorted(i for s in shards for i in s) == list(range(96))
    for r, mine in enumerate(shards):
        part = ctx.basecall_raw([raws[i] for i in mine])
        for k, i in enumerate(mine):
            assert (part.start[k], part.end[k]) == (whole.start[i], whole.end[i])
            assert _same(_per_read(whole, i), _per_read(part, k)), f"rank {r} read {i}"
    # shortest read against the oracle from the raw signal on
    i = int(np.argmin(lens))
    x = hs.prepare_read(raws[i])
    full = oracle.basecall(fm, x, 1.0, False)
    bases, qual = gpu_lib.emit_bases(*whole.read_path(i), fm.nbase)
    assert bases == full["basecall"]
    ctx.close(); m.close()


def test_cfg4_r941_rna002_delta_reverse(gpu_lib, oracle):
    """configs[4]: --model r941_rna002 --reverse --delta 1.0 (LSTM-256 behind three convolutions, stride 5;
    difference_array + /delta instead of med-MAD, reference src/flappie.c:254-259; bases emitted 3'->5' reversed)."""
    fm = FlipflopModel.for_name("r941_rna002", seed=1)
    raws = synthetic_reads(64, 4000, seed=17) + synthetic_reads(8, [1500, 2500, 6000, 9000, 12000, 800, 20000, 3333], seed=18)
    m = Model(fm); ctx = Context(m)
    a = ctx.basecall_raw(raws, delta=1.0, want_trans=True)
    b = ctx.basecall_raw(raws[60:70], delta=1.0, want_trans=True)
    for k in range(10):
        assert _same(_per_read(a, 60 + k), _per_read(b, k))
    for i in (0, 69):
        se = hs.trim_and_segment_raw(raws[i])
        x = (hs.difference_array(raws[i][se[0]:se[1]]) / np.float32(1.0)).astype(np.float32)
        t_o = oracle.transitions(fm, x, 1.0)
        assert np.max(np.abs(a.read_trans(i) - t_o)) < TOL_TRANS
        full = oracle.basecall(fm, x, 1.0, False)
        fwd, fq = gpu_lib.emit_bases(*a.read_path(i), fm.nbase, reverse=False)
        rev, rq = gpu_lib.emit_bases(*a.read_path(i), fm.nbase, reverse=True)
        assert fwd == full["basecall"] and rev == fwd[::-1] and rq == fq[::-1]
    ctx.close(); m.close()


def test_submit_collect_pipeline_two_contexts(gpu_lib):
    """ffb_submit_raw_batch / ffb_collect with two contexts in flight: same bits as the blocking call."""
    import torch
    fm = FlipflopModel.for_name("r10C_pcr", seed=1)
    m = Model(fm)
    batches = [synthetic_reads(24, 3000 + 500 * k, seed=40 + k) for k in range(4)]
    blocking = Context(m)
    want = [blocking.basecall_raw(b) for b in batches]
    ctxs = [Context(m), Context(m)]
    inflight, got = [], []

    def prep(cx, raws):
        lens = np.array([len(r) for r in raws], np.int64)
        off = np.zeros(len(raws) + 1, np.int64); np.cumsum(lens, out=off[1:])
        raw = torch.from_numpy(np.concatenate(raws)).pin_memory().numpy()
        b, o = cx.make_batch(raw, off, 1.0, 0)
        rb, s, e = cx.make_raw_batch(raw, off)
        return rb, b, o, (raw, off, s, e)

    for k, raws in enumerate(batches):
        cx = ctxs[k % 2]
        rb, b, o, keep = prep(cx, raws)
        cx.submit_raw(rb, b)
        inflight.append((cx, b, o, keep, rb))
        if len(inflight) == 2:
            pcx, pb, po, _, _ = inflight.pop(0)
            pcx.collect(pb); got.append(po)
    for pcx, pb, po, _, _ in inflight:
        pcx.collect(pb); got.append(po)
    for w, g in zip(want, got):
        nb = int(w.blk_off[-1]) + w.n_reads
        assert np.array_equal(w.blk_off, g["blk_off"]) and np.array_equal(w.path[:nb], g["path"][:nb])
        assert np.array_equal(w.score, g["score"][:w.n_reads])
    for cx in ctxs + [blocking]:
        cx.close()
    m.close()
