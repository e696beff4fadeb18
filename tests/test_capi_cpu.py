"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every
symbol include/rpf_b200.h declares, fails loudly without a GPU (no CPU
fallback), and the synthetic source is well-formed."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from rpf import pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu():
    import torch
    return torch.cuda.is_available()


def test_header_and_library_agree():
    hdr = open(os.path.join(ROOT, "include", "rpf_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(rpf_[a-z_0-9]+)\s*\(", hdr)))
    assert declared == sorted(pkg.EXPORTED_SYMBOLS)
    lib = pkg.lib()
    for name in declared:
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", pkg.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (rpf_[a-z_0-9]+)", out))
    assert set(declared) <= exported
    assert lib.rpf_abi_version() == 2


def test_every_entry_point_cites_the_reference():
    hdr = open(os.path.join(ROOT, "include", "rpf_b200.h")).read()
    for fn in ("rpf_create", "rpf_destroy", "rpf_reset", "rpf_push", "rpf_finish", "rpf_write_data"):
        i = hdr.index("int " + fn) if ("int " + fn) in hdr else hdr.index("void " + fn)
        assert re.search(r"src/\w+\.(cxx|h):\d+", hdr[max(0, i - 1500):i]), fn


def test_no_oracle_in_product_path():
    """The product may not include, link or call anything under oracle/."""
    pk = os.path.join(ROOT, "rtl-power-fftw_b200")
    for dp, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".cpp", ".py", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in txt.lower() or f in ("__init__.py",), (dp, f)
    out = subprocess.run(["ldd", pkg.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "rpf_ref" not in out


@pytest.mark.skipif(_gpu(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu():
    with pytest.raises(pkg.RpfError) as e:
        pkg.Context(4096)
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)
    exe = os.path.join(ROOT, "rtl-power-fftw_b200", "host", "rpf_power_synth")
    r = subprocess.run([exe, "512", "100"], capture_output=True, text=True)
    assert r.returncode == 7                    # ReturnValue::HardwareError, as the reference would exit
    assert "no usable CUDA device" in r.stderr


def test_create_argument_validation_without_gpu():
    lib = pkg.lib()
    h = C.c_void_p()
    assert lib.rpf_create(None, C.byref(h)) == -1
    cfg = pkg.capi.RpfConfig()
    cfg.N = 511
    assert lib.rpf_create(C.byref(cfg), C.byref(h)) == -1
    assert b"even" in lib.rpf_strerror(None)
    cfg.N = (1 << 22) + 2
    assert lib.rpf_create(C.byref(cfg), C.byref(h)) == -2      # valid for the reference, beyond the CUDA path: loud
    assert b"not supported" in lib.rpf_strerror(None)
    lib.rpf_destroy(None)                                       # NULL allowed


def test_synth_source_is_counter_based():
    a = pkg.synth_fill_host(4096, 0, seed=1, mode=0)
    for off, ln in ((0, 17), (5, 100), (8, 64), (1001, 999)):
        assert np.array_equal(pkg.synth_fill_host(ln, off, seed=1, mode=0), a[off:off + ln])
    assert not np.array_equal(a, pkg.synth_fill_host(4096, 0, seed=2, mode=0))
    g = pkg.synth_fill_host(1 << 20, 0, seed=3, mode=1).astype(np.float64)
    assert abs(g.mean() - 127.4) < 0.2 and abs(g.std() - 30.0) < 0.5
    t = pkg.synth_fill_host(2 * 4096, 0, seed=3, mode=2).astype(np.float64)
    z = (t[0::2] - 127) + 1j * (t[1::2] - 127)
    assert np.argmax(np.abs(np.fft.fft(z))) == 4096 // 8        # +fs/8 tone


def test_shard_ranges_cover_and_are_spectrum_aligned():
    sh = pkg.sharding if hasattr(pkg, "sharding") else __import__("importlib").import_module(
        "rtl-power-fftw_b200.sharding")
    for total, world in ((80000000, 8), (1000003, 8), (7, 8), (100, 3)):
        ranges = [sh.shard_range(total, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == total
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
        assert max(b - a for a, b in ranges) - min(b - a for a, b in ranges) <= 1
    assert sh.shard_byte_range(4096, 80000000, 1, 8) == (2 * 4096 * 10000000, 2 * 4096 * 20000000)


def test_auxdata_text_reader(tmp_path):
    """Window / baseline text files (src/acquisition.cxx:32-156): last number per line, '#' comments,
    blank / non-numeric lines skipped, exactly N values or InvalidInput (exit code 5)."""
    exe = os.path.join(ROOT, "rtl-power-fftw_b200", "host", "rpf_power_synth")
    N = 8
    w = tmp_path / "win.txt"
    w.write_text("# window\n\n  # indented comment\n0.0\n1 0.25\n2 3 0.5\nabc\n0.75\n1.0\n0.75e0\n 5e-1\n0.25\n")
    b = tmp_path / "base.txt"
    b.write_text("".join("%d %.17g\n" % (i, -1.5 + i * 0.1234567890123) for i in range(N)))
    args = [exe, str(N), "1", "0", "1", "1", "0", "5", "0", "0"]
    # window from a file, baseline from stdin (with both from FILES the reference fails to open the second one: it
    # re-opens one std::ifstream, src/acquisition.cxx:62,108,132 -- reproduced below)
    r = subprocess.run(args + ["-w", str(w), "-B", "-", "--dump-aux"], input=b.read_text(), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.split()
    assert lines[0] == "window" and int(lines[1]) == N
    win = np.array(lines[2:2 + N], dtype=np.float64)
    assert np.array_equal(win.astype(np.float32), np.array([0, .25, .5, .75, 1, .75, .5, .25], np.float32))
    assert lines[2 + N] == "baseline" and int(lines[3 + N]) == N
    base = np.array(lines[4 + N:4 + 2 * N], dtype=np.float64)
    assert np.array_equal(base, np.array([-1.5 + i * 0.1234567890123 for i in range(N)]))
    r = subprocess.run(args + ["-w", str(w), "-B", str(b), "--dump-aux"], capture_output=True, text=True)
    assert r.returncode == 5 and "Could not open %s. Quitting." % b in r.stderr
    # both from stdin: first N -> window, next N -> baseline
    both = "".join("%g\n" % v for v in list(range(N)) + [10 + v for v in range(N)])
    r = subprocess.run(args + ["-w", "-", "-B", "-", "--dump-aux"], input=both, capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.split()[2:2 + N] == [str(v) for v in range(N)]
    # wrong count / missing file -> ReturnValue::InvalidInput
    w.write_text("1\n2\n3\n")
    r = subprocess.run(args + ["-w", str(w), "--dump-aux"], capture_output=True, text=True)
    assert r.returncode == 5 and "Expected 8 values, found 3" in r.stderr
    r = subprocess.run(args + ["-B", str(tmp_path / "nope"), "--dump-aux"], capture_output=True, text=True)
    assert r.returncode == 5 and "Could not open" in r.stderr


def test_auxdata_reader_matches_reference_parser(tmp_path):
    """AuxData's text grammar against what the reference's OWN read_inputfile<T>() returns
    (tests/golden/aux_parser.json, generated by tests/golden/make_aux_golden.py from the function
    compiled out of /root/reference): multi-column lines, comments, and the corner case that a
    failed extraction zeroes the kept value ("0.5 abc" -> 0, src/acquisition.cxx:48-56)."""
    import json
    exe = os.path.join(ROOT, "rtl-power-fftw_b200", "host", "rpf_power_synth")
    cases = json.load(open(os.path.join(ROOT, "tests", "golden", "aux_parser.json")))
    assert cases["trailing_garbage_zeroes"]["double"][:2] == ["0", "0"]
    checked = 0
    for name, c in sorted(cases.items()):
        for opt, key, dtype in (("-w", "float", np.float32), ("-B", "double", np.float64)):
            text, want = c["text"], list(c[key])       # float and double may keep different lines (1e40 overflows float)
            if not want:
                continue
            if len(want) % 2:                           # the CLI forces N even (src/params.cxx:152-155)
                text += "\n1\n"
                want.append("1")
            n = len(want)
            f = tmp_path / (name + key + ".txt")
            f.write_bytes(text.encode())
            r = subprocess.run([exe, str(n), "1", "0", "1", "1", "0", "5", "0", "0", opt, str(f), "--dump-aux"],
                               capture_output=True, text=True)
            assert r.returncode == 0, (name, key, r.stderr)
            lines = r.stdout.split()
            at = lines.index("window" if opt == "-w" else "baseline")
            assert int(lines[at + 1]) == n, (name, key)
            got = np.array(lines[at + 2:at + 2 + n], dtype=np.float64).astype(dtype)
            assert np.array_equal(got, np.array(want, dtype=np.float64).astype(dtype)), (name, key)
            checked += 1
    assert checked >= 12
    # where the reference is present (the build container), also compare live
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_read_inputfile")
    if os.path.exists(ref):
        for name, c in cases.items():
            r = subprocess.run([ref, "double"], input=c["text"], capture_output=True, text=True)
            assert r.stdout.split() == c["double"], name


def test_plan_matches_oracle():
    """Plan (host/acquisition.cpp) against the restatement of src/acquisition.cxx:162-198: repeats from the
    integration time, buffer auto-sizing, the list of centre frequencies with overlap.  No GPU involved."""
    import oracle_lib as ol
    exe = os.path.join(ROOT, "rtl-power-fftw_b200", "host", "rpf_power_synth")
    cases = [(512, 2000000, 0.5, 88000000, 108000000, 0), (512, 2400000, 0.1, 88000000, 108000000, 10),
             (4096, 2048000, 1.5, 1419000000, 1421000000, 50), (500, 1000000, 0.01, 100000000, 100500000, 0),
             (1024, 3200000, 2.0, 24000000, 1766000000, 33.3), (512, 2000000, 1e-5, 433000000, 433000001, 0),
             (2048, 2399999, 0.25, 144000000, 146000000, 99)]
    for N, rate, t, f0, f1, ov in cases:
        r = subprocess.run([exe, str(N), "0", "--rate", str(rate), "--integration-time", repr(t), "--freq-range",
                            "%d:%d" % (f0, f1), "--overlap", repr(ov), "--plan-only"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        out = dict((ln.split()[0], ln.split()[1:]) for ln in r.stdout.splitlines())
        rep = ol.oracle().rpf_oracle_plan_repeats(N, rate, t)
        assert int(out["repeats"][0]) == rep
        assert int(out["buf_length"][0]) == ol.oracle().rpf_oracle_plan_buf_length(N, rep, 1638400, 0)
        freqs = ol.oracle_plan_freqs(f0, f1, rate, ov)
        assert int(out["freqs"][0]) == len(freqs) and [int(x) for x in out["freqs"][1:]] == freqs
        assert "Number of averaged spectra: %d" % rep in r.stderr
        # ... and against the reference's OWN Plan::Plan (src/acquisition.cxx:158-198 compiled unmodified into
        # oracle/_ref/ref_rtl_power), where the reference sources are present
        ref_bin = os.path.join(ROOT, "oracle", "_ref", "ref_rtl_power")
        if os.path.exists(ref_bin):
            rr = subprocess.run([ref_bin, "/dev/null", str(N), "0", "--rate", str(rate), "--integration-time", repr(t),
                                 "--freq-range", "%d:%d" % (f0, f1), "--overlap", repr(ov), "--plan-only"],
                                capture_output=True, text=True)
            assert rr.returncode == 0, rr.stderr
            assert rr.stdout == r.stdout, (rr.stdout, r.stdout)
    # without hopping: the single centre frequency (src/acquisition.cxx:195-197)
    r = subprocess.run([exe, "512", "100", "--plan-only"], capture_output=True, text=True)
    assert "freqs 1 1420405752" in r.stdout and "repeats 100" in r.stdout and "buf_length 114688" in r.stdout


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints one JSON line with
    the contract's keys; it times the reference's own worker loop (oracle/_ref) on the host cores."""
    import json
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    if "unavailable" in line:
        return
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["unit"] == "Msamples/s" and line["value"] > 1.0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
