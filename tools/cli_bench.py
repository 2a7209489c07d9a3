"""Files -> fastq throughput of the `flappie` command line (flappie_b200/host/flappie): N single-read signal files on local
disk (tmpfs when available), one process, 1 GPU and all GPUs of the box.  Prints the CLI's own --stats lines.
    python tools/cli_bench.py [n_files=8192] [model=r941_native_gru] [batch=1024]"""
import os, re, shutil, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from flappie_b200.api import Library
from flappie_b200.model import FlipflopModel, synthetic_reads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
name = sys.argv[2] if len(sys.argv) > 2 else "r941_native_gru"
batch = sys.argv[3] if len(sys.argv) > 3 else "1024"
root = tempfile.mkdtemp(prefix="ffb_cli_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
try:
    fm = FlipflopModel.for_name(name, seed=1)
    fm.save_bundle(os.path.join(root, "r941_native.ffbw"))
    rdir = os.path.join(root, "reads"); os.mkdir(rdir)
    t0 = time.time()
    base = synthetic_reads(256, 4000, seed=7)
    rng = np.random.default_rng(1)
    for i in range(n):                                    # 256 distinct squiggles + per-file noise: distinct reads, cheap to make
        (base[i % 256] + rng.normal(0, 0.3, 4000).astype(np.float32)).astype(np.float32).tofile(os.path.join(rdir, f"read_{i:06d}.f32"))
    print(f"# {n} files x 4000 samples written to {rdir} in {time.time() - t0:.1f} s; model {name}, --batch {batch}", flush=True)
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "flappie_b200", "host", "flappie")
    env = dict(os.environ, FLAPPIE_B200_MODELS=root)
    ndev = Library.get().device_count()
    outs = {}
    for devs in ["0"] + (["all"] if ndev > 1 else []) + ["0,0"]:
        for rep in range(2):
            out = os.path.join(root, f"out_{devs.replace(',', '_')}.fastq")
            t0 = time.time()
            r = subprocess.run([exe, "--model", "r941_native", "--devices", devs, "--batch", batch, "--stats", "--output", out, rdir],
                               capture_output=True, text=True, env=env)
            wall = time.time() - t0
            m = re.search(r"stats (\{.*\})", r.stderr)
            print(f"--devices {devs:4s} run {rep}: rc={r.returncode} wall {wall:.2f} s (incl. process start, model upload)  {m.group(1) if m else r.stderr[-300:]}", flush=True)
            if rep == 1:
                for ln in r.stderr.splitlines():
                    if "where the host time went" in ln or "device thread" in ln:
                        print("      " + ln.split(": ", 1)[1], flush=True)
            outs[devs] = open(out, "rb").read() if os.path.exists(out) else b""
    ref = outs["0"]
    for k, v in outs.items():
        print(f"# output of --devices {k}: {len(v)} bytes, identical to --devices 0: {v == ref}")
finally:
    shutil.rmtree(root, ignore_errors=True)
