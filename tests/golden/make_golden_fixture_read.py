"""The reference's own 37 838-sample test read (src/test/raw_signal.crp, BASELINE configs[0] / SURVEY.md 8(d) cfg1) as
a committed fixture, with what the REFERENCE'S OWN OBJECT CODE (oracle/_ref) calls from it.

    python tests/golden/make_golden_fixture_read.py        (build container only: needs /root/reference)

Writes tests/golden/fixture_read.npz:
  raw_adc            int16[37838]  the fixture's samples (ADC counts; pA = (adc + 16) * 1373.41 / 8192,
                                   src/test/test_flappie_signal.c:74-83)
  start, end         kept range after trim_and_segment_raw defaults (src/flappie.c:100-110)
  vit_path           int8[T+1]     decode_crf_flipflop on trans (--viterbi), model r941_native_gru seed 1
  vit_score, vit_bases, vit_quals
  fb_bases, fb_quals, fb_score     default mode (transpost + decode)
  trans_sub          float32[T/32][40]   trans at every 32nd block
The weights are FlipflopModel.for_name("r941_native_gru", seed=1) -- the .mdl files of the reference are git-LFS
pointers.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from flappie_b200.model import FlipflopModel  # noqa: E402
from flappie_b200.signal import trim_and_segment_raw  # noqa: E402
from oracle.pyoracle import Ref  # noqa: E402
from make_golden import read_crp  # noqa: E402

MODEL, SEED = "r941_native_gru", 1


def fixture_pa(raw_adc):
    unit = np.float32(1373.41) / np.float32(8192.0)
    return ((raw_adc.astype(np.float32) + np.float32(16.0)) * unit).astype(np.float32)


def main():
    r = Ref()
    raw = read_crp("/root/reference/src/test/raw_signal.crp")[:, 0]
    adc = raw.astype(np.int16)
    assert np.array_equal(adc.astype(np.float32), raw) and adc.shape[0] == 37838
    pa = fixture_pa(adc)
    s, e = trim_and_segment_raw(pa)
    norm = r.medmad_normalise(pa[s:e])
    fm = FlipflopModel.for_name(MODEL, seed=SEED)
    rm = r.model(fm)
    trans = r.transitions(rm, norm, 1.0)
    vit = r.decode(trans, viterbi_only=True, want_trace=False)
    fb = r.decode(trans, viterbi_only=False, want_trace=False)
    np.savez_compressed(os.path.join(HERE, "fixture_read.npz"), raw_adc=adc, start=s, end=e,
                        vit_path=vit["path"].astype(np.int8), vit_score=np.float32(vit["score"]),
                        vit_bases=np.frombuffer(vit["basecall"].encode(), np.uint8), vit_quals=np.frombuffer(vit["quality"].encode(), np.uint8),
                        fb_bases=np.frombuffer(fb["basecall"].encode(), np.uint8), fb_quals=np.frombuffer(fb["quality"].encode(), np.uint8),
                        fb_score=np.float32(fb["score"]), trans_sub=trans[::32].copy(), buildinfo=r.buildinfo)
    print("blocks", trans.shape[0], "bases", len(vit["basecall"]), len(fb["basecall"]), "kept", s, e, r.buildinfo)


if __name__ == "__main__":
    main()
