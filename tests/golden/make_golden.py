"""Generate the golden vectors under tests/golden/ from the REFERENCE'S OWN OBJECT CODE
(oracle/_ref/libflappie_ref.so, compiled by oracle/Makefile from /root/reference/src).

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

The GPU box has no /root/reference: the -m gpu tests and the oracle tests read only the
.npz files written here.  Everything is seeded; re-running reproduces the files bit for bit
(up to the OpenBLAS build named in oracle/_ref/BUILDINFO).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from flappie_b200.model import KIND_GRU, KIND_LSTM, FlipflopModel, synthetic_reads  # noqa: E402
from oracle.pyoracle import ACT_NONE, ACT_SWISH, ACT_TANH, Ref  # noqa: E402

CASES = [
    # name, kind, size, nbase, model seed, raw length
    ("gru64_5b", KIND_GRU, 64, 5, 21, 1400),
    ("gru96_4b", KIND_GRU, 96, 4, 22, 1300),
    ("lstm96_4b", KIND_LSTM, 96, 4, 23, 1500),
]


def read_crp(path, limit=None):
    """reference src/test/flappie_util.c:79-132 text format: 'nr\\tnc' then one line of
    hex floats per column."""
    with open(path) as fh:
        nr, nc = (int(x) for x in fh.readline().split())
        cols = []
        for i, line in enumerate(fh):
            if limit is not None and i >= limit:
                break
            cols.append([float.fromhex(t) for t in line.split()])
    return np.asarray(cols, np.float32)


def main():
    r = Ref()
    # ---- 1. signal fixtures of the reference's own test-suite (src/test/*.crp) ------------
    # pins medmad_normalise_array / trim on real data: test_flappie_signal.c:67-111.
    tdir = "/root/reference/src/test"
    raw = read_crp(os.path.join(tdir, "raw_signal.crp"))[:, 0]
    trimmed = read_crp(os.path.join(tdir, "trimmed_signal.crp"))[:, 0]
    normalised = read_crp(os.path.join(tdir, "normalised_signal.crp"))[:, 0]
    # keep the fixture small: the first 6000 raw samples and what the reference code itself
    # produces from them (the full-length check against the .crp files runs in
    # tests/test_oracle.py whenever /root/reference is present)
    unit = np.float32(1373.41) / np.float32(8192.0)
    raw_pa = ((raw + np.float32(16.0)) * unit).astype(np.float32)      # test_flappie_signal.c:74-83
    head = raw_pa[:6000].copy()
    from flappie_b200.signal import trim_and_segment_raw
    se = trim_and_segment_raw(head)
    head_norm = r.medmad_normalise(head[se[0]:se[1]])
    np.savez_compressed(os.path.join(HERE, "signal_fixture.npz"), raw_pa_head=head, start=se[0], end=se[1],
                        normalised_head=head_norm, full_raw_len=raw.shape[0], full_trim_len=trimmed.shape[0],
                        full_norm_first=normalised[:64], full_norm_last=normalised[-64:])

    # ---- 2. convolution incl. the right-edge behaviour -------------------------------------
    rng = np.random.default_rng(5)
    conv = {}
    for stride, winlen, nf, nfilter, T in ((2, 19, 1, 8, 100), (2, 19, 1, 8, 101), (5, 19, 16, 8, 100),
                                           (5, 19, 16, 8, 103), (1, 5, 4, 16, 57), (2, 19, 1, 8, 3790),
                                           (5, 19, 16, 4, 3790), (3, 19, 1, 4, 60), (1, 5, 1, 4, 19)):
        x = rng.normal(size=(T, nf)).astype(np.float32)
        W = (rng.normal(size=(nfilter, winlen, nf)) * 0.3).astype(np.float32)
        b = rng.normal(size=(nfilter,)).astype(np.float32)
        key = f"s{stride}_w{winlen}_nf{nf}_f{nfilter}_T{T}"
        conv[key + "_x"], conv[key + "_W"], conv[key + "_b"] = x, W, b
        conv[key + "_y"] = r.convolution(x, W, b, stride, ACT_NONE)
    np.savez_compressed(os.path.join(HERE, "conv_golden.npz"), **conv)

    # ---- 3. decode pieces on random scores --------------------------------------------------
    dec = {}
    for nbase, T in ((4, 257), (5, 130)):
        nr = 2 * nbase * (nbase + 1)
        trans = (rng.normal(size=(T, nr)) * 2).astype(np.float32)
        score, path, qpath = r.viterbi(trans)
        tpost = r.transpost(trans, True)
        s2, p2, q2 = r.viterbi(tpost)
        trace = r.trace(np.exp(tpost).astype(np.float32))
        k = f"b{nbase}"
        dec.update({k + "_trans": trans, k + "_score": np.float32(score), k + "_path": path, k + "_qpath": qpath,
                    k + "_tpost": tpost, k + "_post_score": np.float32(s2), k + "_post_path": p2, k + "_trace": trace})
    np.savez_compressed(os.path.join(HERE, "decode_golden.npz"), **dec)

    # ---- 4. whole networks ---------------------------------------------------------------------
    from flappie_b200.signal import prepare_read
    for name, kind, size, nbase, seed, rawlen in CASES:
        fm = FlipflopModel.synthetic(kind, size, nbase, seed=seed)
        sig = prepare_read(synthetic_reads(1, rawlen, seed=seed + 100)[0])
        rm = r.model(fm)
        trans, conv_o, layers = r.transitions(rm, sig, 1.0, want_layers=True)
        dv = r.decode(trans, True, True)
        df = r.decode(trans, False, True)
        np.savez_compressed(
            os.path.join(HERE, f"net_{name}.npz"), kind=kind, size=size, nbase=nbase, seed=seed, signal=sig,
            conv=conv_o.astype(np.float16), layer1=layers[0].astype(np.float16), layer5=layers[4], trans=trans,
            vit_path=dv["path"], vit_score=np.float32(dv["score"]), vit_bases=dv["basecall"], vit_qual=dv["quality"],
            fb_path=df["path"], fb_score=np.float32(df["score"]), fb_bases=df["basecall"], fb_qual=df["quality"],
            fb_trace=df["trace"].astype(np.uint8), buildinfo=r.buildinfo)
        print(name, "blocks", trans.shape[0], "bases", len(dv["basecall"]), len(df["basecall"]))


if __name__ == "__main__":
    main()
