#!/usr/bin/env python
"""bench.py -- raw-signal samples/sec basecalled (BASELINE.json metric) on N B200s.

A "step" is one pass of the hot path (network forward + flip-flop decode, the default
forward-backward + Viterbi mode of the reference CLI) over one batch of synthetic reads:
BASELINE.json configs[1] = 1024 synthetic 4000-sample reads, r941_native, 1 GPU.
Per-GPU work is fixed as N grows (reads shard embarrassingly, no collective on the data
path): "scaling": "weak".

  value : whole-job samples/s with the normalised signals already resident in HBM
          (CUDA events on the library's stream around exactly K x ffb_forward()).
  e2e   : same metric through the reference-facing C-ABI calls ffb_submit_raw_batch() /
          ffb_collect() with HOST buffers: H2D of the RAW samples from pinned memory, trimming +
          normalisation + network + decoding + base/quality emission on the device, D2H of the
          called bases, quality characters and scores -- all inside the timed region.
  extra : the other BASELINE configs (r941_native LSTM-384 x1024, r941_5mC x4096, r10C_pcr mixed
          1 k-50 k, r941_rna002 --delta --reverse) with value / e2e / roofline.frac each, and the
          STRONG-scaling set: a fixed 24576-read configs[3] workload dealt over the N ranks by
          flappie_b200.shard.shard_reads.
  roofline     : the recurrent-layer kernel (dominant), algorithmic flops / CUDA-event time.
  cpu_baseline : the reference's own code (oracle/_ref, OpenBLAS, 1 thread per process,
                 one process per core as its README recommends) on a bounded sample.

`--impl reference` times only the CPU reference arm.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

RAW_SAMPLES = 4000
MODEL_CHOICES = {
    # name -> (MODEL_TABLE key, description)
    "r941_native_gru": ("r941_native_gru", "r941_native as in the north star / flappie 1.x: conv(19,s2,tanh) + 5 grumod S=256, 4 bases"),
    "r941_native": ("r941_native", "r941_native @4de542f: 3 conv(swish) + 5 LSTM S=384, stride 5"),
    "r941_5mC": ("r941_5mC", "conv + 5 grumod S=256, 5 bases"),
    "r941_rna002": ("r941_rna002", "3 conv + 5 LSTM S=256"),
    "r10C_pcr": ("r10C_pcr", "alias: conv + 5 grumod S=256, 4 bases"),
    "r103_native": ("r103_native", "3 conv + 5 LSTM S=512 (16-CTA clusters, lo weight plane in shared memory)"),
    "rle_r941_native": ("rle_r941_native", "runnie: 3 conv + 5 LSTM S=256 + run-length head"),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


def make_workload(model_key: str, n_reads: int, raw_len: int, seed: int):
    """Seeded weights + reads; host signal prep (trim + med-MAD, reference flappie.c:251-259)
    is done ONCE here, outside every timed region: it precedes the hot path."""
    from flappie_b200.model import FlipflopModel, synthetic_reads
    from flappie_b200.signal import prepare_read
    fm = FlipflopModel.for_name(model_key, seed=1)
    raws = synthetic_reads(n_reads, raw_len, seed=seed)
    reads = [prepare_read(r) for r in raws]
    assert all(r is not None for r in reads)
    return fm, reads, raws


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0=None, t1=None):
        """Summarise the samples taken inside [t0, t1] (perf_counter); nvidia-smi needs a moment to start on an 8-GPU
        box, so the sampler is started before the warm-up and, if the timed window itself caught no sample, the samples
        taken under the warm-up load (same kernels) stand in -- the summary says which."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        window = "timed region"
        rows = [r for t, r in self.rows if (t0 is None or t >= t0) and (t1 is None or t <= t1 + 0.15)]
        if not rows:
            rows = [r for t, r in self.rows]
            window = "warm-up + timed region (no sample fell inside the timed region)"
        sm, mx, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
                power.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for k, nm in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "power_w": float(np.median(power)) if power else None, "samples": len(sm), "window": window}


# ------------------------------------------------------------------------------------------
def _ref_worker(args):
    """One process = one core, OpenBLAS pinned to one thread (reference README.md:66-67,80-83)."""
    model_key, read_arrays = args
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    from flappie_b200.model import FlipflopModel
    from oracle.pyoracle import Ref
    r = Ref()
    fm = FlipflopModel.for_name(model_key, seed=1)
    rm = r.model(fm)
    t0 = time.perf_counter()
    nb = 0
    for sig in read_arrays:
        out = r.basecall(rm, sig, 1.0, False)
        nb += len(out["basecall"]) if out else 0
    return time.perf_counter() - t0, nb


def cpu_reference_rate(model_key: str, reads, cores: int, reads_per_core: int):
    """samples/s of the reference's CPU path on `cores` processes over a bounded sample."""
    import multiprocessing as mp
    from oracle import pyoracle
    if not pyoracle.have_ref():
        raise RuntimeError("oracle/_ref/libflappie_ref.so missing (build it where /root/reference exists)")
    n = min(len(reads), cores * reads_per_core)
    cores = min(cores, n)
    chunks = [reads[i::cores][:reads_per_core] for i in range(cores)]
    nread = sum(len(c) for c in chunks)
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_ref_worker, [(model_key, c) for c in chunks])
    wall = time.perf_counter() - t0
    busy = max(r[0] for r in res)
    rate = nread * RAW_SAMPLES / busy
    info = pyoracle.Ref().buildinfo
    return dict(value=rate, unit="samples/s", cores=cores, kind="reference",
                sample=f"{nread} of the workload's reads ({reads_per_core}/core), fwd-bwd + Viterbi + trace as calculate_post; "
                       f"slowest worker {busy:.1f}s, wall {wall:.1f}s incl. process start; {info}")


# ------------------------------------------------------------------------------------------
def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the very workload of the b200 arm (rank 0's reads): the config printed below is computed from it, not asserted
    fm, reads, _ = make_workload(MODEL_CHOICES[a.model][0], a.reads, RAW_SAMPLES, seed=7)
    tot_blocks = sum(max(fm.nblock(len(r)), 0) for r in reads)
    cores = len(os.sched_getaffinity(0))
    per_core = max(1, a.ref_reads_per_core)
    rates = []
    for step in range(a.warmup + a.steps):
        cb = cpu_reference_rate(MODEL_CHOICES[a.model][0], reads, cores, per_core)
        if step >= a.warmup:
            rates.append(cb)
    val = float(np.mean([c["value"] for c in rates]))
    cb = dict(rates[-1]); cb["value"] = val
    nread = min(len(reads), cores * per_core)
    line = {
        "impl": "reference", "metric": "raw-signal samples/sec basecalled", "value": val, "unit": "samples/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * nread * RAW_SAMPLES / val, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (the reference's own fp32 / OpenBLAS code)", "data": "synthetic",
        # the SAME config as the b200 arm measures; the CPU arm times a bounded sample of it and extrapolates the rate
        # (a rate metric): which sample is in cpu_baseline.sample
        "config": build_config(a, a.gpus, tot_blocks),
        "reference_sample": f"each step = {nread} of the {a.reads} reads ({per_core} per core on {cores} host cores), rate extrapolated",
        "cpu_baseline": cb,
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def mixed_lengths(n, lo, hi, seed):
    """configs[3]: lengths log-uniform in [lo, hi] samples"""
    rng = np.random.default_rng(seed)
    return [int(x) for x in np.exp(rng.uniform(np.log(lo), np.log(hi), n))]


def run_extras(a, api, lib, peaks, world, rank, local, barrier, max_over_ranks):
    """The other BASELINE configs, one short measurement each (1 warm-up + 2 timed steps): `value` = K x ffb_forward with the
    batch resident after ONE ffb_upload_raw (CUDA events), `e2e` = the blocking C-ABI call ffb_basecall_raw_batch from pinned
    host memory with device-side emission (wall clock), roofline.frac of the recurrent kernel from ffb_forward_timed."""
    import ctypes
    import torch
    from flappie_b200.model import FlipflopModel, synthetic_reads
    from flappie_b200.shard import shard_reads
    out = []
    cases = [
        ("configs[1b] r941_native @4de542f (LSTM-384, stride 5), 1024 x 4000", "r941_native", [RAW_SAMPLES] * 1024, {}),
        ("configs[2] r941_5mC (GRU-256, 5 bases), 4096 x 4000", "r941_5mC", [RAW_SAMPLES] * 4096, {}),
        ("configs[3] r10C_pcr (GRU-256), 3072 reads log-uniform 1 k-50 k samples per GPU", "r10C_pcr", mixed_lengths(3072, 1000, 50000, 100 + rank), {}),
        ("configs[4] r941_rna002 (LSTM-256) --delta 1.0 --reverse, 1024 x 4000", "r941_rna002", [RAW_SAMPLES] * 1024, {"delta": 1.0, "reverse": True}),
    ]
    # the strong-scaling set: ONE fixed configs[3] workload, dealt over the ranks by shard_reads (LPT on length), every
    # rank runs its shard in batches of at most 3072 reads
    strong_lens = mixed_lengths(a.strong_reads, 1000, 50000, 4242)
    mine = shard_reads(strong_lens, world)[rank]
    cases.append((f"configs[3] STRONG scaling: fixed {a.strong_reads}-read set (log-uniform 1 k-50 k) sharded over {world} GPU(s) by shard_reads",
                  "r10C_pcr", [strong_lens[i] for i in mine], {"strong": True}))
    for title, name, lens, opt in cases:
        fm = FlipflopModel.for_name(name, seed=1)
        model = api.Model(fm, device=local)
        stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
        ctx = api.Context(model, stream=stream.cuda_stream)
        flags = api.FLAG_REVERSE if opt.get("reverse") else 0
        chunks = [lens[i::(len(lens) + 3071) // 3072] for i in range((len(lens) + 3071) // 3072)] if len(lens) > 4096 else [lens]
        work = []
        for ci, cl in enumerate(chunks):
            raws = synthetic_reads(len(cl), cl, seed=1000 * ci + 11 + rank)
            off = np.zeros(len(cl) + 1, np.int64); np.cumsum([len(r) for r in raws], out=off[1:])
            raw = torch.empty(int(off[-1]), dtype=torch.float32).pin_memory().numpy(); raw[:] = np.concatenate(raws)
            blocks = sum(max(fm.nblock(int(x)), 0) for x in cl)
            o = {"blk_off": np.zeros(len(cl) + 1, np.int64), "score": np.zeros(len(cl), np.float32),
                 "bases": torch.empty(blocks + len(cl), dtype=torch.uint8).pin_memory().numpy(),
                 "quals": torch.empty(blocks + len(cl), dtype=torch.uint8).pin_memory().numpy(), "nbases": np.zeros(len(cl), np.int32)}
            b_, o = ctx.make_batch(raw, off, 1.0, flags, o, emit=True, want_path=False)
            rb_, st_, en_ = ctx.make_raw_batch(raw, off, delta=opt.get("delta", 0.0))
            work.append((rb_, b_, o, raw, (off, st_, en_)))      # everything the C structs point at stays referenced here
            del raws
        samples = int(sum(lens))
        K = 2
        # e2e: blocking raw call per chunk, host buffers in, bases out
        for rb_, b_, o, _, _ in work[:1]:
            ctx._check(lib.lib.ffb_basecall_raw_batch(ctx.handle, ctypes.byref(rb_), ctypes.byref(b_)), "warm-up")
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            for rb_, b_, o, _, _ in work:
                ctx._check(lib.lib.ffb_basecall_raw_batch(ctx.handle, ctypes.byref(rb_), ctypes.byref(b_)), "ffb_basecall_raw_batch")
        barrier()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / K
        entry = {"config": title, "model": name, "reads_this_rank": len(lens), "samples_this_rank": samples}
        if opt.get("strong"):
            total = int(sum(strong_lens))
            entry.update({"scaling": "strong", "n_gpus": world, "value": total / (e2e_ms * 1e-3), "unit": "samples/s",
                          "ms_per_pass": e2e_ms, "what": "whole fixed set / max-over-ranks wall time of the blocking C-ABI calls (host buffers in, bases out)",
                          "shard_imbalance": float(max_over_ranks(float(samples)) * world / total)})
        else:
            # device-resident value + roofline of the recurrent kernel (single chunk by construction)
            rb_, b_, o, _, _ = work[0]
            ctx._check(lib.lib.ffb_upload_raw(ctx.handle, ctypes.byref(rb_), ctypes.byref(b_)), "ffb_upload_raw")
            ctx.forward(); barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            sampler = ClockSampler(local); sampler.start()
            tw0 = time.perf_counter()
            e0.record(stream)
            for _ in range(K):
                ctx.forward()
            e1.record(stream)
            barrier()
            clocks = sampler.stop(tw0, time.perf_counter())
            dev_ms = max_over_ranks(e0.elapsed_time(e1)) / K
            groups = ctx.forward_timed()
            T = ctx.total_blocks()
            S, G = fm.size, fm.ngate
            fl = 2.0 * T * S * G * S + ((4 * 2.0 * T * S * S + 2.0 * T * S * fm.nparam) / 5.0 if (fm.kind == 0 and S in (256, 384)) else 0.0)
            ach = fl / (groups["rnn"] / 5.0 * 1e-3) / 1e12
            entry.update({"scaling": "weak", "n_gpus": world, "value": samples * world / (dev_ms * 1e-3), "unit": "samples/s", "ms_per_step": dev_ms,
                          "e2e": {"value": samples * world / (e2e_ms * 1e-3), "unit": "samples/s", "ms_per_step": e2e_ms,
                                  "what": "blocking ffb_basecall_raw_batch, one context (no overlap of host and device)"},
                          "blocks_this_rank": int(T), "roofline_frac": ach / peaks["tf_sustained"], "rnn_ms_per_launch": groups["rnn"] / 5.0,
                          "step_breakdown_ms": groups, "clocks": clocks})
        out.append(entry)
        ctx.close(); model.close()
        del work, ctx, model
        torch.cuda.empty_cache()
    return out


def build_config(a, world, tot_blocks):
    return {"workload": f"{a.reads} synthetic {RAW_SAMPLES}-sample reads per GPU ({RAW_SAMPLES - 210} after the default trim), "
                        f"{a.model}: {MODEL_CHOICES[a.model][1]}; random-init weights; "
                        f"{'--viterbi' if a.viterbi_only else 'forward-backward + Viterbi (CLI default)'}",
            "blocks_per_gpu": int(tot_blocks) if tot_blocks is not None else None, "reads_per_gpu": a.reads,
            "l2": "working set per step (Xin + activations, ~10 GB) far exceeds the 126 MB L2; no flush needed",
            "schedule": "layer l+1's input GEMM streamed behind layer l's recurrence (PDL); GRU: its z-gate third computed by the recurrence itself"
                        if os.environ.get("FFB_NO_STREAM_GEMM") is None else "sequential kernels",
            "signal_prep": "value: normalised signal resident in HBM (prepared once, outside the timed region); "
                           "e2e: RAW signal from pinned host memory, trimming + med-MAD normalisation on the device inside the timed "
                           "region, two batches in flight (ffb_submit_raw_batch / ffb_collect on two contexts), bases + quality "
                           "characters emitted on the device and copied back, wall clock",
            "parallelism": f"read-shard x{world}, no collective"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="r941_native_gru", choices=sorted(MODEL_CHOICES))
    ap.add_argument("--reads", type=int, default=1024, help="reads per GPU per step")
    ap.add_argument("--viterbi-only", action="store_true")
    ap.add_argument("--fp32-simt", action="store_true", help="force the fp32 CUDA-core GEMM / recurrence kernels")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-reads-per-core", type=int, default=6)
    ap.add_argument("--ref-reads-total", type=int, default=256)
    ap.add_argument("--no-extra", action="store_true", help="skip the `extra` configs (A/B runs)")
    ap.add_argument("--strong-reads", type=int, default=24576,
                    help="reads of the fixed configs[3] strong-scaling set (3072 per GPU at N=8: a batch large enough that its "
                         "longest read -- 24 895 dependent steps per layer -- does not dominate)")
    a = ap.parse_args()

    if a.impl == "reference":
        run_reference_arm(a)
        return

    import torch
    import torch.distributed as dist
    from flappie_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    peaks = load_peaks()

    model_key = MODEL_CHOICES[a.model][0]
    # weak scaling: every rank gets its own `reads` reads (different seed per rank)
    fm, reads, raws = make_workload(model_key, a.reads, RAW_SAMPLES, seed=7 + rank)
    n = len(reads)
    lens = np.array([len(r) for r in reads], np.int64)
    sig_off_np = np.zeros(n + 1, np.int64); np.cumsum(lens, out=sig_off_np[1:])
    sig_pinned = torch.empty(int(sig_off_np[-1]), dtype=torch.float32).pin_memory()
    sig_pinned.numpy()[:] = np.concatenate(reads)
    signal_np = sig_pinned.numpy()

    lib = api.Library.get()
    model = api.Model(fm, device=local)
    stream = torch.cuda.Stream()          # a real (non-default) stream: torch.cuda.Event sees only the stream it records on
    torch.cuda.set_stream(stream)
    ctx = api.Context(model, stream=stream.cuda_stream)
    flags = (api.FLAG_VITERBI_ONLY if a.viterbi_only else 0) | (api.FLAG_FP32_SIMT if a.fp32_simt else 0)
    tot_blocks = sum(max(fm.nblock(int(x)), 0) for x in lens)
    out = {
        "blk_off": np.zeros(n + 1, np.int64),
        "path": torch.empty(tot_blocks + n, dtype=torch.int32).pin_memory().numpy(),
        "qpath": torch.empty(tot_blocks + n, dtype=torch.float32).pin_memory().numpy(),
        "score": torch.empty(n, dtype=torch.float32).pin_memory().numpy(),
    }
    batch, out = ctx.make_batch(signal_np, sig_off_np, 1.0, flags, out)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident arm: inputs in HBM, K x forward ----
    ctx.upload(batch)
    sampler = ClockSampler(local); sampler.start()
    for _ in range(a.warmup):
        ctx.forward()
    barrier()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.perf_counter()
    e0.record(stream)
    for _ in range(a.steps):
        ctx.forward()
    e1.record(stream)
    barrier()
    clocks = sampler.stop(tw0, time.perf_counter())
    launches = ctx.launch_count() - l0
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    samples_per_step = n * RAW_SAMPLES * world
    value = samples_per_step * a.steps / (dev_ms * 1e-3)
    ctx.download(batch)                    # path / qpath of this arm: the e2e arm's device-emitted bases are checked against them

    # ---- end-to-end arm: RAW host buffers through the C ABI (ffb_basecall_raw_batch): H2D of the raw samples,
    #      trimming + normalisation + network + decode on the device, D2H of path/qpath/score, all inside the timed region ----
    import ctypes
    raw_lens = np.array([len(r) for r in raws], np.int64)
    raw_off_np = np.zeros(n + 1, np.int64); np.cumsum(raw_lens, out=raw_off_np[1:])
    raw_pinned = torch.empty(int(raw_off_np[-1]), dtype=torch.float32).pin_memory()
    raw_pinned.numpy()[:] = np.concatenate(raws)
    raw_np = raw_pinned.numpy()
    raw_blocks = sum(max(fm.nblock(int(x)), 0) for x in raw_lens)       # outputs sized for the untrimmed lengths
    # two contexts on two streams: while the device works on batch i the host trims / plans / uploads batch i+1
    # (ffb_submit_raw_batch / ffb_collect); every step still pays its own H2D, device prep and D2H.  The step ends with the
    # called bases and quality characters in host memory (emitted on the device): path / qpath stay in HBM.
    stream2 = torch.cuda.Stream()
    ctx2 = api.Context(model, stream=stream2.cuda_stream)

    def pinned(shape, dtype):
        return torch.empty(shape, dtype=dtype).pin_memory().numpy()

    pipe = []
    for cx in (ctx, ctx2):
        o = {"blk_off": np.zeros(n + 1, np.int64), "score": pinned(n, torch.float32), "bases": pinned(raw_blocks + n, torch.uint8),
             "quals": pinned(raw_blocks + n, torch.uint8), "nbases": pinned(n, torch.int32)}
        b_, o = cx.make_batch(raw_np, raw_off_np, 1.0, flags, o, emit=True, want_path=False)
        rb_, rstart, rend = cx.make_raw_batch(raw_np, raw_off_np)
        pipe.append((cx, rb_, b_, o))
    for cx, rb_, b_, o in pipe:                       # warm-up: workspaces of both contexts
        for _ in range(min(a.warmup, 2)):
            cx.submit_raw(rb_, b_); cx.collect(b_)
    # the device-prepared, device-emitted path must reproduce the host-prepared one: same blocks, and the bases that
    # ffb_emit_bases makes of the device-resident arm's path
    assert np.array_equal(pipe[0][3]["blk_off"], out["blk_off"]) and np.array_equal(pipe[1][3]["blk_off"], out["blk_off"])
    for i in (0, n // 2, n - 1):
        s0 = int(out["blk_off"][i]) + i
        nb_i = int(pipe[0][3]["nbases"][i])
        want_b, want_q = lib.emit_bases(out["path"][s0:s0 + int(out["blk_off"][i + 1] - out["blk_off"][i]) + 1],
                                        out["qpath"][s0:s0 + int(out["blk_off"][i + 1] - out["blk_off"][i]) + 1], fm.nbase)
        assert pipe[0][3]["bases"][s0:s0 + nb_i].tobytes().decode() == want_b and pipe[1][3]["quals"][s0:s0 + nb_i].tobytes().decode() == want_q
    barrier()
    t0 = time.perf_counter()
    for i in range(a.steps):
        cx, rb_, b_, _ = pipe[i % 2]
        cx.submit_raw(rb_, b_)
        if i > 0:
            pcx, _, pb_, _ = pipe[(i - 1) % 2]
            pcx.collect(pb_)
    cx, _, b_, _ = pipe[(a.steps - 1) % 2]
    cx.collect(b_)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    e2e_value = samples_per_step * a.steps / (e2e_ms * 1e-3)
    out_raw = pipe[0][3]
    h2d = int(raw_np.nbytes + 3 * raw_off_np.nbytes + 4 * n)
    d2h = int(out_raw["blk_off"][-1] + n) * 2 + 8 * n + 16 * n     # bases + quality chars of the blocks produced, counts, scores, trim bounds

    # ---- per-kernel-group timing for the roofline (up to three extra passes of the same step, sequential schedule,
    #      CUDA events on the library's stream around each kernel group, averaged; not part of `value`) ----
    passes = [ctx.forward_timed() for _ in range(max(1, min(a.steps, 3)))]
    groups = {k: float(np.mean([p_[k] for p_ in passes])) for k in passes[0]}
    S, G, T = fm.size, fm.ngate, tot_blocks
    tensor_path = not a.fp32_simt and fm.size in (256, 384)
    fused = tensor_path and fm.kind == 0
    rnn_flops_sw = 2.0 * T * S * G * S                   # one layer's h_{t-1} * sW, all reads (algorithmic)
    # GRU: four of the five launches also compute the z-gate third of the next layer's input projection (2*T*S*S each) and the
    # fifth the output layer (2*T*S*nparam) in the same MMAs -- useful flops of the path that the input / output GEMMs no
    # longer do; average per launch
    rnn_flops = rnn_flops_sw + ((4 * 2.0 * T * S * S + 2.0 * T * S * fm.nparam) / 5.0 if fused else 0.0)
    rnn_ms_per_launch = groups["rnn"] / 5.0
    achieved_tf = rnn_flops / (rnn_ms_per_launch * 1e-3) / 1e12
    # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this kernel, from the committed `ncu --set full` capture of
    # this very command (a number taken under the profiler cannot be re-measured inside a timed run): read from the file
    # tools/ncu_summary.py wrote, never a constant in this script
    traffic, traffic_src = None, None
    tj = os.path.join(ROOT, "profiles", "r02_rnn_tc_traffic.json")
    if tensor_path and a.model == "r941_native_gru" and a.reads == 1024 and os.path.exists(tj):
        tdoc = json.load(open(tj))
        traffic, traffic_src = float(tdoc["dram_bytes_per_launch"]), f"profiles/r02_rnn_tc_traffic.json ({tdoc.get('source', 'ncu --set full')})"
    roofline = {"kernel": "rnn_tc_kernel (recurrent layer: h*sW on tcgen05 + gates, 5 launches/step)" if tensor_path
                          else "rnn_layer_kernel (fp32 CUDA-core cluster kernel, 5 launches/step)",
                "bound": "tensor", "achieved": achieved_tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                "frac": achieved_tf / peaks["tf_sustained"], "traffic": traffic, "traffic_source": traffic_src,
                "achieved_sW_only": rnn_flops_sw / (rnn_ms_per_launch * 1e-3) / 1e12,
                "frac_sW_only": rnn_flops_sw / (rnn_ms_per_launch * 1e-3) / 1e12 / peaks["tf_sustained"],
                "timed": "CUDA events around each launch in the SEQUENTIAL schedule (ffb_forward_timed, extra passes after the timed "
                         "region); `value` runs the streamed schedule, where the same launches overlap the next layer's input GEMM",
                "peak_source": f"{peaks['src']} bf16 dense sustained (kernel timed inside a long step)",
                "algorithmic_bytes_per_launch": float(T) * (4 * G * S + 4 * S + (4 * S if fused else 0)),
                "note": "algorithmic flops per launch = 2*blocks*S*G*S (h*sW) + for the GRU the rows that ride in the fourth quarter of "
                        "each M=128 tile: the z-gate third of the next layer's input projection (2*blocks*S*S, four launches) or the "
                        "output layer (fifth launch) -- work the input / output GEMMs no longer do; `*_sW_only` leaves them out (the "
                        "round-1 definition).  The fp32-faithful fp16 hi/lo split issues 3 MMAs per product; the layer is T dependent "
                        "steps of ~2.5 us each, i.e. bound by the latency of the step chain, not by the pipe (DESIGN.md 4.2)",
                "step_breakdown_ms": groups}

    line = {
        "metric": "raw-signal samples/sec basecalled", "value": value, "unit": "samples/s",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dev_ms / a.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (GEMM operands split into fp16 hi + lo, 3 tcgen05 MMAs per product, fp32 accumulate; gates ex2/rcp.approx)",
        "data": "synthetic",
        "config": build_config(a, world, tot_blocks),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / a.steps},
        "gpu_launches": int(launches),
        "roofline": roofline,
    }
    ctx2.close(); ctx.close(); model.close()
    del ctx2, ctx, model
    if not a.no_extra:
        line["extra"] = run_extras(a, api, lib, peaks, world, rank, local, barrier, max_over_ranks)
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            cores = len(os.sched_getaffinity(0))
            line["cpu_baseline"] = cpu_reference_rate(model_key, reads, cores, a.ref_reads_per_core)
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "samples/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
