"""Cycle breakdown of one recurrent step (control thread + one gate warp of cluster 0 / CTA 0 /
group 0).  Needs a library built with FFB_EXTRA_NVCC_FLAGS=-DFFB_RNN_PROFILE."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from flappie_b200.api import Context, Library, Model
from flappie_b200.model import FlipflopModel, synthetic_reads
from flappie_b200.signal import prepare_read

name = sys.argv[1] if len(sys.argv) > 1 else "r941_native_gru"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
L = Library.get().lib
L.ffb_test_rnn_prof.restype = ctypes.c_int
L.ffb_test_rnn_prof.argtypes = [ctypes.POINTER(ctypes.c_uint64), ctypes.c_int]
fm = FlipflopModel.for_name(name, seed=1)
reads = [prepare_read(r) for r in synthetic_reads(n, 4000, seed=7)]
m = Model(fm); ctx = Context(m)
ctx.basecall(reads, viterbi_only=True)
out = (ctypes.c_uint64 * 16)()
L.ffb_test_gemm_prof.restype = ctypes.c_int; L.ffb_test_gemm_prof.argtypes = [ctypes.POINTER(ctypes.c_uint64), ctypes.c_int]; L.ffb_test_gemm_prof(None, 1)
assert L.ffb_test_rnn_prof(out, 1) == 1, "library not built with -DFFB_RNN_PROFILE"
ctx.basecall(reads, viterbi_only=True)
L.ffb_test_rnn_prof(out, 0)
steps = 5 * fm.nblock(len(reads[0]))
names = {0: "ctl wait h_full", 1: "ctl issue MMAs", 2: "ctl wait acc_full", 3: "ctl arrive peers + wait staged", 4: "ctl bulk store + wait_group",
         5: "ctl wait h_empty", 6: "ctl issue multicast", 8: "gate prefetch Xin", 9: "gate wait acc_full", 10: "gate tmem ld", 11: "gate cells + stage"}
tot_c = sum(out[i] for i in range(7)); tot_g = sum(out[i] for i in range(8, 12))
for i, nm in names.items():
    print(f"{nm:34s} {out[i] / steps:9.1f} clk/step")
print(f"control total {tot_c / steps:.1f}  gate total {tot_g / steps:.1f}  ({steps} steps)")
L.ffb_test_gemm_prof.restype = ctypes.c_int
L.ffb_test_gemm_prof.argtypes = [ctypes.POINTER(ctypes.c_uint64), ctypes.c_int]
g = (ctypes.c_uint64 * 16)()
L.ffb_test_gemm_prof(g, 0)
nt = max(int(g[10]), 1)
gn = {0: "prod loop overhead", 1: "prod dependency poll + fence", 2: "prod wait empty stage", 3: "prod issue TMA", 4: "mma wait acc_empty",
      5: "mma wait full stage", 6: "mma issue + commit", 8: "epi wait acc_full", 9: "epi drain + store"}
print(f"GEMM (CTA 0, {nt} tiles since reset, all launches):")
for i, nm in gn.items():
    print(f"{nm:34s} {g[i] / nt:9.1f} clk/tile")

if g[12]:
    t0 = out[12]
    print(f"timeline (us, relative to the start of recurrent layer 4 of 5): recurrence ends {(out[13]-t0)/1e3:.0f}; "
          f"streamed GEMM of layer 5: CTA 0 starts {(int(g[12])-t0)/1e3:.0f}, last CTA starts {(int(g[14])-t0)/1e3:.0f}, "
          f"CTA 0 ends {(int(g[13])-t0)/1e3:.0f} after {int(g[15])-1} tiles")
