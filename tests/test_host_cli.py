"""The C99 host side (flappie_b200/host): record writers against the reference's own fprintf_format
(src/flappie_output.c, compiled into oracle/_ref), weight bundles, raw-signal readers, and -- on the GPU -- the
`flappie` command line end to end against the Python host over the same C ABI."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from flappie_b200.model import KIND_GRU, FlipflopModel, synthetic_reads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "flappie_b200", "host")
GOLD = os.path.join(ROOT, "tests", "golden")


class ReadResult(ctypes.Structure):
    _fields_ = [("score", ctypes.c_float), ("n", ctypes.c_size_t), ("start", ctypes.c_size_t), ("end", ctypes.c_size_t),
                ("basecall", ctypes.c_char_p), ("quality", ctypes.c_char_p), ("basecall_length", ctypes.c_size_t),
                ("nblock", ctypes.c_size_t)]


@pytest.fixture(scope="module")
def host():
    path = os.path.join(HOST, "libffb_host.so")
    if not os.path.exists(path):
        from flappie_b200.build import build_host
        build_host()
    L = ctypes.CDLL(path)
    L.ffb_get_outformat.restype = ctypes.c_int; L.ffb_get_outformat.argtypes = [ctypes.c_char_p]
    L.ffb_outformat_string.restype = ctypes.c_char_p; L.ffb_outformat_string.argtypes = [ctypes.c_int]
    L.ffb_fprintf_read.restype = None
    L.ffb_fprintf_read.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_bool,
                                   ctypes.c_char_p, ctypes.POINTER(ReadResult)]
    L.ffb_read_raw_file.restype = ctypes.c_long
    L.ffb_read_raw_file.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.POINTER(ctypes.c_float))]
    return L


def _libc():
    libc = ctypes.CDLL(None)
    libc.fopen.restype = ctypes.c_void_p; libc.fopen.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
    libc.fclose.argtypes = [ctypes.c_void_p]
    libc.free.argtypes = [ctypes.c_void_p]
    return libc


def _write_mine(host, fmt, path, uuid, readname, uuid_primary, prefix, score, nblock, basecall, quality, n, start, end):
    libc = _libc()
    fp = libc.fopen(path.encode(), b"a")
    res = ReadResult(score, n, start, end, basecall.encode(), quality.encode() if quality is not None else None,
                     len(basecall), nblock)
    host.ffb_fprintf_read(host.ffb_get_outformat(fmt.encode()), fp, uuid.encode(), readname.encode(), uuid_primary,
                          prefix.encode(), ctypes.byref(res))
    libc.fclose(fp)


RECORDS = [
    ("0f776a08-1101-41d4-8097-89136494a46e", "read_ch1_file0.fast5", True, "", -1234.5678, 1895, "ACGTTGCA" * 30, "5678:;<=" * 30, 4000, 200, 3990),
    ("uuid-2", "b.fast5", False, "run7_", -0.03125, 7, "ACGTZ", "!~+5I", 100, 0, 90),
    ("u3", "c.f32", True, "p", -3e5, 24895, "A", "#", 50000, 1300, 49690),
]


@pytest.mark.parametrize("fmt", ["fasta", "fastq", "sam"])
def test_record_writers_match_reference_bytes(host, ref, tmp_path, fmt):
    mine, theirs = str(tmp_path / "mine.txt"), str(tmp_path / "ref.txt")
    for rec in RECORDS:
        _write_mine(host, fmt, mine, *rec)
        ref.format_record(fmt, theirs, *rec)
    a, b = open(mine, "rb").read(), open(theirs, "rb").read()
    assert a == b and len(a) > 0
    if fmt == "sam":            # the reference prints sequence and quality twice (src/flappie_output.c:126-131)
        assert a.count(b"ACGTZ") == 2


def test_outformat_names(host):
    for i, nm in enumerate([b"fasta", b"fastq", b"sam"]):
        assert host.ffb_get_outformat(nm) == i and host.ffb_outformat_string(i) == nm
    assert host.ffb_get_outformat(b"bam") == 3 and host.ffb_outformat_string(3) is None


def test_fastq_without_quality_prints_nothing(host, tmp_path):
    p = str(tmp_path / "x.txt")
    _write_mine(host, "fastq", p, "u", "r", True, "", -1.0, 10, "ACGT", None, 100, 0, 90)
    assert open(p, "rb").read() == b""          # warnx + return, src/flappie_output.c:108-111


def test_raw_readers(host, tmp_path):
    x = np.random.default_rng(0).normal(90, 12, 1234).astype(np.float32)
    f32 = tmp_path / "a.f32"; x.tofile(f32)
    crp = tmp_path / "a.crp"
    with open(crp, "w") as fh:                  # write_flappie_matrix_to_handle layout, src/test/flappie_util.c:30-55
        fh.write(f"1\t{x.shape[0]}\n")
        for v in x:
            fh.write(float(v).hex() + "\n")
    libc = _libc()
    for path in (f32, crp):
        ptr = ctypes.POINTER(ctypes.c_float)()
        n = host.ffb_read_raw_file(str(path).encode(), ctypes.byref(ptr))
        assert n == x.shape[0]
        assert np.array_equal(np.ctypeslib.as_array(ptr, shape=(n,)), x)
        libc.free(ptr)
    ptr = ctypes.POINTER(ctypes.c_float)()
    assert host.ffb_read_raw_file(b"/nonexistent/x.f32", ctypes.byref(ptr)) == -1
    assert host.ffb_read_raw_file(b"whatever.fast5", ctypes.byref(ptr)) == -2     # needs libhdf5


def test_reference_crp_fixture_is_readable(host):
    path = "/root/reference/src/test/raw_signal.crp"
    if not os.path.exists(path):
        pytest.skip("reference fixtures not mounted")
    ptr = ctypes.POINTER(ctypes.c_float)()
    n = host.ffb_read_raw_file(path.encode(), ctypes.byref(ptr))
    assert n == 37838
    g = np.load(os.path.join(GOLD, "signal_fixture.npz"))
    unit = np.float32(1373.41) / np.float32(8192.0)
    head = ((np.ctypeslib.as_array(ptr, shape=(n,))[:g["raw_pa_head"].shape[0]] + np.float32(16.0)) * unit).astype(np.float32)
    assert np.array_equal(head, g["raw_pa_head"])
    _libc().free(ptr)


def test_cli_refuses_without_gpu_or_weights(tmp_path):
    exe = os.path.join(HOST, "flappie")
    r = subprocess.run([exe, "--model", "help"], capture_output=True, text=True)
    assert r.returncode == 0 and "r941_native" in r.stdout and "(default)" in r.stdout
    r = subprocess.run([exe, "--format", "bam", "x.f32"], capture_output=True, text=True)
    assert r.returncode != 0 and "Unrecognised output format" in r.stderr
    r = subprocess.run([exe, "--model", "nosuch", "x.f32"], capture_output=True, text=True)
    assert r.returncode != 0 and "Invalid Flappie model" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("fmt,extra", [("fastq", []), ("fasta", ["--viterbi"]), ("sam", ["--reverse", "--no-uuid", "--prefix", "x_"])])
def test_cli_end_to_end(gpu_lib, ref, tmp_path, fmt, extra):
    """flappie <dir of .f32 reads>: records equal, byte for byte, the reference's writers fed with the Python
    host's results over the same C ABI (same kernels => same bits)."""
    from flappie_b200.api import Context, Model
    fm = FlipflopModel.synthetic(KIND_GRU, 96, 4, seed=3, name="r941_native")
    fm.save_bundle(str(tmp_path / "r941_native.ffbw"))
    lens = [4000, 2500, 6000, 150, 4000, 3000, 90]
    raws = synthetic_reads(len(lens), lens, seed=31)
    rdir = tmp_path / "reads"; rdir.mkdir()
    names = [f"read_{i:02d}.f32" for i in range(len(lens))]
    for nm, r in zip(names, raws):
        r.tofile(rdir / nm)
    out = tmp_path / f"calls.{fmt}"
    env = dict(os.environ, FLAPPIE_B200_MODELS=str(tmp_path))
    r = subprocess.run([os.path.join(HOST, "flappie"), "--format", fmt, "--batch", "3", "--output", str(out)] + extra + [str(rdir)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr
    assert r.stderr.count("No basecall returned") == 2          # the 150- and 90-sample reads trim to nothing

    m = Model(fm); ctx = Context(m)
    want = str(tmp_path / "want.txt")
    viterbi, reverse, uuid_primary = "--viterbi" in extra, "--reverse" in extra, "--no-uuid" not in extra
    prefix = "x_" if "--prefix" in extra else ""
    # same batches as the command line made (batch composition does not change bits, but keep it honest)
    for b0 in range(0, len(lens), 3):
        res = ctx.basecall_raw(raws[b0:b0 + 3], viterbi_only=viterbi)
        for k in range(res.n_reads):
            i = b0 + k
            if res.nblock(k) == 0:
                continue
            bases, qual = gpu_lib.emit_bases(*res.read_path(k), fm.nbase, reverse=reverse)
            ref.format_record(fmt, want, names[i][:-4], names[i], uuid_primary, prefix, float(res.score[k]), res.nblock(k), bases, qual,
                              lens[i], int(res.start[k]), int(res.end[k]))
    assert open(out, "rb").read() == open(want, "rb").read()
    ctx.close(); m.close()


@pytest.mark.gpu
@pytest.mark.parametrize("extra", [[], ["--viterbi"]])
def test_runnie_cli_end_to_end(gpu_lib, tmp_path, extra):
    """runnie <dir of .f32 reads>: the `.run` text (reference src/runnie.c:277-310) equals the Python host's results
    over the same C ABI, printed with the same format."""
    from flappie_b200.api import Context, Model
    from flappie_b200.model import KIND_LSTM
    fm = FlipflopModel.synthetic(KIND_LSTM, 96, 4, seed=3, name="rle_r941_native")
    fm.head = "runlength"
    fm.save_bundle(str(tmp_path / "rle_r941_native.ffbw"))
    lens = [4000, 2500, 6000, 150, 3000]
    raws = synthetic_reads(len(lens), lens, seed=33)
    rdir = tmp_path / "reads"; rdir.mkdir()
    names = [f"read_{i:02d}.f32" for i in range(len(lens))]
    for nm, r in zip(names, raws):
        r.tofile(rdir / nm)
    out = tmp_path / "calls.run"
    env = dict(os.environ, FLAPPIE_B200_MODELS=str(tmp_path))
    r = subprocess.run([os.path.join(HOST, "runnie"), "--batch", "2", "--output", str(out)] + extra + [str(rdir)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr
    m = Model(fm); ctx = Context(m)
    want = []
    for b0 in range(0, len(lens), 2):
        res = ctx.basecall_raw(raws[b0:b0 + 2], viterbi_only="--viterbi" in extra)
        for k in range(res.n_reads):
            if res.nblock(k) == 0:
                continue
            st, rle = res.read_rle(k)
            bases, shape, scale, dwell = gpu_lib.emit_runs(st, rle)
            want.append("# %s\n" % names[b0 + k][:-4])
            want += ["%s\t%f\t%f\t%d\n" % (c, float(sh), float(sc), int(d)) for c, sh, sc, d in zip(bases, shape, scale, dwell)]
    assert open(out).read() == "".join(want) and len(want) > 50
    ctx.close(); m.close()


def test_weight_bundle_roundtrip(host, tmp_path):
    """FlipflopModel.save_bundle -> ffb_bundle_load (C): the `_Mat` images arrive in the reference's struct order with
    their padded columns (no device needed)."""
    from flappie_b200.model import KIND_LSTM, Mat

    class Bundle(ctypes.Structure):
        _fields_ = [("kind", ctypes.c_int), ("nconv", ctypes.c_int), ("nmat", ctypes.c_int), ("stride", ctypes.c_int * 3),
                    ("mats", ctypes.POINTER(Mat))]

    host.ffb_bundle_load.restype = ctypes.c_int
    host.ffb_bundle_load.argtypes = [ctypes.c_char_p, ctypes.POINTER(Bundle)]
    host.ffb_bundle_free.restype = None
    host.ffb_bundle_free.argtypes = [ctypes.POINTER(Bundle)]
    for kind, nmat, strides, head in ((KIND_GRU, 19, [2, 0, 0], "flipflop"), (KIND_LSTM, 23, [1, 1, 5], "flipflop"),
                                      (KIND_LSTM, 23, [1, 1, 5], "runlength")):
        fm = FlipflopModel.synthetic(kind, 64, 4, seed=9)
        fm.head = head
        path = str(tmp_path / f"m{kind}{head}.ffbw")
        fm.save_bundle(path)
        b = Bundle()
        assert host.ffb_bundle_load(path.encode(), ctypes.byref(b)) == 0
        assert b.kind == (2 if head == "runlength" else kind) and b.nmat == nmat and list(b.stride) == strides
        mats, keep = fm.to_mat_bundle()
        for i, m in enumerate(mats):
            got = b.mats[i]
            assert (got.nr, got.nc, got.stride) == (m.nr, m.nc, m.stride)
            n = m.nc * m.stride
            assert np.array_equal(np.ctypeslib.as_array(got.data, shape=(n,)), np.ctypeslib.as_array(m.data, shape=(n,)))
        host.ffb_bundle_free(ctypes.byref(b))
    bad = tmp_path / "bad.ffbw"
    bad.write_bytes(b"not a bundle")
    assert host.ffb_bundle_load(str(bad).encode(), ctypes.byref(Bundle())) != 0


def _write_reads(tmp_path, lens, seed):
    raws = synthetic_reads(len(lens), lens, seed=seed)
    rdir = tmp_path / "reads"; rdir.mkdir()
    names = [f"read_{i:03d}.f32" for i in range(len(lens))]
    for nm, r in zip(names, raws):
        r.tofile(rdir / nm)
    return rdir, names, raws


@pytest.mark.gpu
@pytest.mark.parametrize("devices", ["0,0", "all", "0,0,0"])
def test_cli_sharded_over_devices_prints_the_single_device_bytes(gpu_lib, tmp_path, devices):
    """--devices: reads of a window are dealt longest-first to the least-loaded device, every device runs its own batch,
    records come out in INPUT order -- byte for byte what one device prints (reference semantics: one record per file in
    the order of the command line, src/flappie.c:364-385).  "0,0" = two independent pipelines on one GPU, so the path runs
    on a 1-GPU box too; "all" uses every GPU of the box."""
    fm = FlipflopModel.synthetic(KIND_GRU, 96, 4, seed=3, name="r941_native")
    fm.save_bundle(str(tmp_path / "r941_native.ffbw"))
    rng = np.random.default_rng(8)
    lens = [int(x) for x in np.exp(rng.uniform(np.log(600), np.log(9000), 61))] + [150, 90]   # ragged, two rejects
    rng.shuffle(lens)
    rdir, names, raws = _write_reads(tmp_path, lens, 37)
    env = dict(os.environ, FLAPPIE_B200_MODELS=str(tmp_path))
    outs = {}
    for tag, opts in (("one", ["--device", "0", "--batch", "64"]), ("many", ["--devices", devices, "--batch", "9"])):
        out = tmp_path / f"{tag}.fastq"
        r = subprocess.run([os.path.join(HOST, "flappie"), "--output", str(out), "--stats"] + opts + [str(rdir)],
                           capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        assert r.stderr.count("No basecall returned") == 2 and "samples_per_s" in r.stderr
        outs[tag] = open(out, "rb").read()
    assert outs["one"] == outs["many"]
    heads = [ln.split()[0][1:] for ln in outs["one"].decode().splitlines() if ln.startswith("@read_")]
    assert heads == [n[:-4] for n, ln in zip(names, lens) if ln > 400]          # input order, rejects dropped


@pytest.mark.gpu
def test_cli_trace_sink(gpu_lib, tmp_path):
    """--trace: the u8 trace of every read (reference src/flappie.c:299-300, written to HDF5 by fast5_interface.c:126-143)
    in the flat record file of ffb_host.h; equals trace_from_posterior over the same C ABI."""
    from flappie_b200.api import Context, Model
    fm = FlipflopModel.synthetic(KIND_GRU, 96, 4, seed=3, name="r941_native")
    fm.save_bundle(str(tmp_path / "r941_native.ffbw"))
    lens = [3000, 1500, 4200]
    rdir, names, raws = _write_reads(tmp_path, lens, 41)
    env = dict(os.environ, FLAPPIE_B200_MODELS=str(tmp_path))
    tr = tmp_path / "trace.bin"
    r = subprocess.run([os.path.join(HOST, "flappie"), "--trace", str(tr), "--output", str(tmp_path / "o.fastq"), str(rdir)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr
    m = Model(fm); ctx = Context(m)
    res = ctx.basecall_raw(raws, want_trace=True)
    blob = open(tr, "rb").read()
    off = 0
    for i, nm in enumerate(names):
        assert blob[off:off + 4] == b"FFBT"
        ln = int(np.frombuffer(blob, np.uint32, 1, off + 4)[0]); off += 8
        assert blob[off:off + ln].decode() == nm[:-4]; off += ln
        nrow = int(np.frombuffer(blob[off:off + 8], np.uint64)[0]); ns = int(np.frombuffer(blob[off + 8:off + 12], np.uint32)[0]); off += 12
        assert nrow == res.nblock(i) + 1 and ns == 8
        got = np.frombuffer(blob, np.uint8, nrow * ns, off).reshape(nrow, ns); off += nrow * ns
        assert np.array_equal(got, res.read_trace(i))
        assert abs(int(got[1:].sum(axis=1).mean()) - 255) <= 3           # posterior mass per block ~ 1
    assert off == len(blob)
    ctx.close(); m.close()
