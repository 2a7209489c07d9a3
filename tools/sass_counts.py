"""profiles/r02_sass_counts.txt: which Blackwell-only instructions the built library really contains, per kernel
(cuobjdump -sass of flappie_b200/csrc/libflappie_b200.so; no GPU needed).
UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk (TMA bulk copy),
UTCBAR = tcgen05.commit, SYNCS = mbarrier operations, MUFU = special-function unit."""
import collections
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "flappie_b200", "csrc", "libflappie_b200.so")
KEYS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "MUFU"]
sass = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
counts = collections.OrderedDict()
fn = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        counts[fn] = collections.Counter()
        continue
    if fn:
        for k in KEYS:
            if re.search(r"\b" + k + r"\b|\b" + k + r"\.", line):
                counts[fn][k] += 1
names = subprocess.run(["/usr/local/cuda/bin/cu++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
nv = subprocess.run(["/usr/local/cuda/bin/nvcc", "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-1]
print(f"# {time.strftime('%Y-%m-%dT%H:%M:%SZ', time.gmtime())}  {nv}  libflappie_b200.so ({os.path.getsize(LIB)} bytes), sm_100a")
print(f"# {'kernel':100s} " + " ".join(f"{k:>8s}" for k in KEYS))
tot = collections.Counter()
for (fn, c), nm in zip(counts.items(), names):
    tot.update(c)
    if any(c[k] for k in KEYS[:6]):
        nm = (nm.split(">(")[0] + ">") if ">(" in nm else re.sub(r"\(.*", "", nm)
        nm = nm.replace("(int)", "").replace("(bool)", "").replace("void ", "")
        print(f"{nm[:102]:102s} " + " ".join(f"{c[k]:8d}" for k in KEYS))
print(f"{'TOTAL (all ' + str(len(counts)) + ' kernels)':102s} " + " ".join(f"{tot[k]:8d}" for k in KEYS))
