"""Golden vectors for the run-length ("runnie") head and for quantile / MAD, from the REFERENCE'S OWN OBJECT CODE
(oracle/_ref/libflappie_ref.so).  Run where /root/reference exists:

    python tests/golden/make_golden_rle.py

Writes tests/golden/rle_golden.npz (seeded; re-running reproduces it)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle.pyoracle import Ref  # noqa: E402


def main():
    r = Ref()
    rng = np.random.default_rng(41)
    out = {}
    T, S, nr = 180, 24, 40
    h = rng.uniform(-1, 1, (T, S)).astype(np.float32)
    W = (rng.normal(size=(nr, S)) * 0.4).astype(np.float32)
    b = (rng.normal(size=nr) * 0.2).astype(np.float32)
    out.update(h=h, W=W, b=b)
    for temp in (1.0, 0.7):
        p = r.globalnorm_runlength(h, W, b, temp)                 # globalnorm_runlengthV2, layers.c:1326-1358
        out[f"param_t{temp}"] = p
    p = out["param_t1.0"]
    s, path = r.rle_viterbi(p)                                    # decode_crf_runlength, decode.c:901-984
    out.update(vit_score=np.float32(s), vit_path=path.astype(np.int32))
    post = r.rle_transpost(p)                                     # transpost_crf_runlength, decode.c:1013-1159
    s2, path2 = r.rle_viterbi(post)
    out.update(post=post, post_score=np.float32(s2), post_path=path2.astype(np.int32))
    q = rng.integers(-2, 3, size=(120, 40)).astype(np.float32)    # ties
    out.update(tie_param=q, tie_path=r.rle_viterbi(q)[1].astype(np.int32))
    # quantilef / madf / medmad_normalise_array (util.c:100-212)
    x = np.round(rng.normal(90, 12, 257), 1).astype(np.float32)
    out.update(sig=x, sig_q=np.array([r.quantile(x, p_) for p_ in (0.0, 0.05, 0.3, 0.5, 0.77, 1.0)], np.float32),
               sig_mad=np.float32(r.mad(x)), sig_norm=r.medmad_normalise(x))
    np.savez_compressed(os.path.join(HERE, "rle_golden.npz"), **out)
    print("wrote rle_golden.npz", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
