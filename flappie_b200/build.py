"""Build libflappie_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m flappie_b200.build [--force]

The library lands in flappie_b200/csrc/ so that it travels with the repo snapshot to
the GPU box (built artefacts are git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(CSRC, "libflappie_b200.so")
SOURCES = ["conv.cu", "signal.cu", "gemm.cu", "gemm_tc.cu", "rnn.cu", "rnn_tc.cu", "decode.cu", "emit.cu", "rle.cu", "api.cu", "testhooks.cu"]
HEADERS = ["ffb_common.cuh", "tc_common.cuh", os.path.join("..", "..", "include", "flappie_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=default", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    extra = os.environ.get("FFB_EXTRA_NVCC_FLAGS", "").split()   # e.g. -DFFB_RNN_PROFILE (then use --force)
    os.makedirs(OBJ, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    jobs = []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s.replace(".cu", ".o"))
        if force or _stale(obj, [src] + hdrs):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose and r.stderr.strip():
            print(r.stderr, file=sys.stderr)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=6) as ex:
        list(ex.map(compile_one, jobs))
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in srcs]
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                     "-Xlinker", "-Bsymbolic"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    build_host(force or bool(jobs))
    return LIB


HOST = os.path.join(HERE, "host")
HOST_SOURCES = ["ffb_output.c", "ffb_weights.c", "ffb_rawio.c", "ffb_shard.c"]
HOST_BIN = os.path.join(HOST, "flappie")
HOST_LIB = os.path.join(HOST, "libffb_host.so")


def build_host(force: bool = False) -> None:
    """The C99 host side: the `flappie` command line and, for the tests, the same objects as a shared library."""
    cc = shutil.which("gcc") or "gcc"
    srcs = [os.path.join(HOST, s) for s in HOST_SOURCES]
    deps = srcs + [os.path.join(HOST, "ffb_host.h"), os.path.join(HOST, "flappie_main.c"), LIB,
                   os.path.join(HERE, "..", "include", "flappie_b200.h")]
    flags = ["-std=c99", "-O2", "-Wall", "-Wextra", "-D_POSIX_C_SOURCE=200809L", "-fPIC"]
    link = ["-L" + CSRC, "-lflappie_b200", "-Wl,-rpath,$ORIGIN/../csrc", "-lm", "-lpthread"]
    if force or _stale(HOST_BIN, deps):
        r = subprocess.run([cc] + flags + ["-o", HOST_BIN, os.path.join(HOST, "flappie_main.c")] + srcs + link,
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"gcc failed for the flappie command line:\n{r.stdout}\n{r.stderr}")
    runnie = os.path.join(HOST, "runnie")
    if force or _stale(runnie, deps):
        r = subprocess.run([cc] + flags + ["-DFFB_RUNNIE", "-o", runnie, os.path.join(HOST, "flappie_main.c")] + srcs + link,
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"gcc failed for the runnie command line:\n{r.stdout}\n{r.stderr}")
    if force or _stale(HOST_LIB, deps):
        r = subprocess.run([cc] + flags + ["-shared", "-o", HOST_LIB] + srcs + link, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"gcc failed for libffb_host.so:\n{r.stdout}\n{r.stderr}")


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
