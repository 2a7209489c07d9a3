"""Manual GPU debugging aid: per-stage deviation of the CUDA path from the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from flappie_b200.api import Context, Model, Library
from flappie_b200.model import KIND_GRU, KIND_LSTM, FlipflopModel, synthetic_reads
from flappie_b200.signal import prepare_read
from oracle.pyoracle import Oracle

def main():
    orc = Oracle()
    for kind, size, nbase in ((KIND_GRU, 64, 4), (KIND_GRU, 256, 4), (KIND_LSTM, 96, 4), (KIND_LSTM, 256, 4), (KIND_LSTM, 384, 4)):
        fm = FlipflopModel.synthetic(kind, size, nbase, seed=1)
        reads = [prepare_read(r) for r in synthetic_reads(3, 1200, seed=3)]
        reads[1] = reads[1][:700]
        try:
            m = Model(fm); ctx = Context(m)
            res = ctx.basecall(reads, viterbi_only=True, want_trans=True, keep_layers=True)
        except Exception as e:
            print("model", kind, size, "FAILED", e); continue
        conv_g = ctx.fetch_layer(0); layers_g = [ctx.fetch_layer(1 + l) for l in range(5)]
        for i, sig in enumerate(reads):
            trans_o, conv_o, layers_o = orc.transitions(fm, sig, 1.0, want_layers=True)
            b0, b1 = int(res.blk_off[i]), int(res.blk_off[i + 1])
            print(f"kind {kind} S {size} read {i} T {b1-b0}: conv {np.abs(conv_g[b0:b1]-conv_o).max():.2e}",
                  "layers", " ".join(f"{np.abs(layers_g[l][b0:b1]-layers_o[l]).max():.2e}" for l in range(5)),
                  f"trans {np.abs(res.read_trans(i)-trans_o).max():.2e}")
            if i == 0:
                d = np.abs(layers_g[0][b0:b1]-layers_o[0])
                print("   layer0 err by time (first 4, last 4):", d.max(axis=1)[:4], d.max(axis=1)[-4:], " by hidden max idx", int(d.max(axis=0).argmax()))
                # affine check: logZ
                print("   logZ gpu", ctx.fetch_logz(3))
        ctx.close(); m.close()

if __name__ == "__main__":
    main()
