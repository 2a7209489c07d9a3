"""One small batch through every kernel of a model: the workload compute-sanitizer is pointed at
(tools/sanitize.sh).  64 reads of mixed length so that ragged groups, the edge plan and the frozen-state
path all run; prints a checksum so repeated runs can be compared."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zlib
import numpy as np
from flappie_b200.api import Context, Model
from flappie_b200.model import FlipflopModel, synthetic_reads

name = sys.argv[1] if len(sys.argv) > 1 else "r941_native_gru"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
nsamp = int(sys.argv[3]) if len(sys.argv) > 3 else 900
rng = np.random.default_rng(3)
raws = [r[: int(rng.integers(nsamp // 2, nsamp + 1))] for r in synthetic_reads(n, nsamp, seed=11)]
fm = FlipflopModel.for_name(name, seed=1)
m = Model(fm); ctx = Context(m)
res = ctx.basecall_raw(raws, viterbi_only=False)
crc = 0
for i in range(n):
    p, q = res.read_path(i)
    crc = zlib.crc32(np.ascontiguousarray(p).tobytes(), crc)
print(f"{name}: {n} reads, blocks {ctx.total_blocks()}, launches {ctx.launch_count()}, path crc {crc:08x}")
