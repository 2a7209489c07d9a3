"""Link-level proof of the drop-in boundary and BASELINE configs[0].

oracle/_ref/flappie_linkproof is the reference's own src/flappie.c -- main(), argp options, calculate_post, UNMODIFIED
-- compiled against the reference's own headers together with its util / trimming / output sources, but WITHOUT
layers.c decode.c nnfeatures.c flappie_matrix.c networks.c: every symbol of those five files resolves into
libflappie_b200.so (oracle/Makefile, target linkproof; fast5 input replaced by a .f32 reader because libhdf5 is not in
this image).  Built in the container that has /root/reference; the GPU box runs the prebuilt binary."""
import os
import subprocess

import numpy as np
import pytest

from flappie_b200 import api
from flappie_b200.model import FlipflopModel, synthetic_reads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "flappie_linkproof")
HOST = os.path.join(ROOT, "flappie_b200", "host")
GOLD = os.path.join(ROOT, "tests", "golden")
REPLACED = {"calculate_transitions", "transpost_crf_flipflop", "decode_crf_flipflop", "trace_from_posterior",
            "exp_activation_inplace", "nbase_from_flipflop_nparam", "change_positions", "free_flappie_matrix",
            "free_flappie_imatrix", "get_flappie_model_type", "flappie_model_string", "flappie_model_description"}


def _need_exe():
    if not os.path.exists(EXE):
        if os.path.isdir("/root/reference/src"):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "linkproof"], check=True, capture_output=True)
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/flappie_linkproof not built (reference sources absent)")


def fixture_pa():
    g = np.load(os.path.join(GOLD, "fixture_read.npz"))
    unit = np.float32(1373.41) / np.float32(8192.0)          # test_flappie_signal.c:74-83
    return g, ((g["raw_adc"].astype(np.float32) + np.float32(16.0)) * unit).astype(np.float32)


def test_reference_main_links_against_this_library_only():
    _need_exe()
    out = subprocess.run(["nm", "-D", "--undefined-only", EXE], capture_output=True, text=True, check=True).stdout
    und = {ln.split()[-1] for ln in out.splitlines() if ln.strip()}
    ours = {s for s in und if "@" not in s and not s.startswith("_")}
    assert ours == REPLACED, ours ^ REPLACED                  # exactly what flappie.c takes from the five replaced files
    assert ours <= set(api.EXPORTS)
    needed = subprocess.run(["readelf", "-d", EXE], capture_output=True, text=True, check=True).stdout
    libs = [ln.split("[")[1].rstrip("]") for ln in needed.splitlines() if "NEEDED" in ln]
    assert "libflappie_b200.so" in libs and not any("flappie_ref" in x or "blas" in x.lower() for x in libs), libs
    r = subprocess.run([EXE, "--model", "help"], capture_output=True, text=True)       # the reference's own option table
    assert r.returncode == 0 and "r941_native" in r.stdout and "r103_native" in r.stdout


def test_fixture_read_golden_is_consistent(oracle):
    """the committed golden of the reference's 37 838-sample test read: sizes, kept range, and the oracle restatement's
    Viterbi path on it (scalar fp32 vs the reference's OpenBLAS build: the same calls wherever the float noise allows)"""
    g, pa = fixture_pa()
    assert pa.shape[0] == 37838 and (int(g["start"]), int(g["end"])) == (200, 37790)   # flappie_common.c:13-81 defaults
    T = g["vit_path"].shape[0] - 1
    assert T == (37590 + 1) // 2 and g["trans_sub"].shape == ((T + 31) // 32, 40)
    from flappie_b200.signal import prepare_read
    fm = FlipflopModel.for_name("r941_native_gru", seed=1)
    trans = oracle.transitions(fm, prepare_read(pa), 1.0)
    assert np.max(np.abs(trans[::32] - g["trans_sub"])) < 5e-5
    _, path, _ = oracle.viterbi(trans)
    assert np.mean(path.astype(np.int8) != g["vit_path"]) < 2e-3


@pytest.mark.gpu
def test_fixture_read_through_the_per_read_dropins(gpu_lib):
    """BASELINE configs[0] as SURVEY.md 8(d) defines it: the reference's own test read through calculate_transitions ->
    decode_crf_flipflop (the calls of src/flappie.c:261-283, --viterbi) against the reference's object code: path
    identical block for block, trans within the north-star 1e-4."""
    from flappie_b200.api import Model, RawTable
    from flappie_b200.signal import prepare_read
    g, pa = fixture_pa()
    fm = FlipflopModel.for_name("r941_native_gru", seed=1)
    m = Model(fm)
    m.register("r941_native")
    sig = prepare_read(pa)                                    # bit-exact host restatement of trim + med-MAD (test_oracle.py)
    trans = gpu_lib.calculate_transitions(sig, 1.0, 0)
    assert trans.shape == (g["vit_path"].shape[0] - 1, 40)
    d = float(np.max(np.abs(trans[::32] - g["trans_sub"])))
    assert d < 1e-4, d
    score, path, qpath = gpu_lib.decode_crf_flipflop(trans)
    nd = int(np.count_nonzero(path.astype(np.int8) != g["vit_path"]))
    assert nd == 0, f"{nd} of {path.shape[0]} blocks differ from the reference's Viterbi path"
    assert abs(score - float(g["vit_score"])) < 2e-3 * abs(float(g["vit_score"])) + 1.0
    bases, quals = gpu_lib.emit_bases(path, qpath, 4)
    assert bases == g["vit_bases"].tobytes().decode()
    m.close()


@pytest.mark.gpu
@pytest.mark.parametrize("extra", [[], ["--viterbi"], ["--reverse", "--format", "sam"]])
def test_reference_main_on_this_library_equals_the_b200_cli(gpu_lib, tmp_path, extra):
    """The reference's main() + calculate_post (one read per call, host trimming / normalisation by the reference's own
    util.c) running on libflappie_b200.so writes, byte for byte, what the batched `flappie` command line of this repo
    writes for the same files -- the fixture read included -- and calls the bases the reference's CPU code called."""
    _need_exe()
    g, pa = fixture_pa()
    fm = FlipflopModel.for_name("r941_native_gru", seed=1)
    fm.name = "r941_native"
    fm.save_bundle(str(tmp_path / "r941_native.ffbw"))
    lens = [4000, 2500, 6000, 150, 3000]
    raws = synthetic_reads(len(lens), lens, seed=31) + [pa]
    rdir = tmp_path / "reads"; rdir.mkdir()
    files = []
    for i, r in enumerate(raws):
        np.asarray(r, np.float32).tofile(rdir / f"read_{i:02d}.f32")
        files.append(str(rdir / f"read_{i:02d}.f32"))       # named one by one: the reference globs "<dir>/*.fast5" only
    # default mode: the batched path skips the -logZ/T shift of trans (the posteriors are invariant under it, api.cu) -- which
    # moves their last bits and with them the sixth decimal of the printed score.  FFB_ALWAYS_LOGZ=1 keeps the shift, as the
    # per-read calls of the reference main do: then the two outputs agree byte for byte in every mode.
    env = dict(os.environ, FLAPPIE_B200_MODELS=str(tmp_path), FFB_ALWAYS_LOGZ="1")
    outs = []
    for exe, more in ((EXE, []), (os.path.join(HOST, "flappie"), ["--batch", "4"])):
        out = tmp_path / (os.path.basename(exe) + ".out")
        r = subprocess.run([exe, "--model", "r941_native", "--output", str(out)] + more + extra + files,
                           capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(open(out, "rb").read())
    assert outs[0] == outs[1] and len(outs[0]) > 10000
    if "--format" not in extra:
        rec = outs[0].decode().split("@read_05")[1].split("\n")           # fastq record of the fixture read
        want = g["vit_bases" if "--viterbi" in extra else "fb_bases"].tobytes().decode()
        same = sum(a == b for a, b in zip(rec[1], want))
        assert len(rec[1]) == len(want) and same >= len(want) - 2, (len(rec[1]), len(want), same)
