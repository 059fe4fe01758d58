"""The `alice-codec` CLI (alice-codec_b200/cli.py) against the reference's CLI semantics (src/bin/main.rs:36-196)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = [sys.executable, os.path.join(ROOT, "alice-codec_b200", "cli.py")]


def run(*args):
    return subprocess.run(CLI + list(args), capture_output=True, text=True)


def test_info_and_errors_need_no_device(tmp_path):
    rgb = O.generate(O.G0, 8, 8, 2)
    alc = O.encode(rgb, 8, 8, 2, 90, 1)
    p = tmp_path / "a.alc"
    p.write_bytes(alc)
    r = run("info", str(p))
    assert r.returncode == 0
    lines = r.stdout.splitlines()
    assert lines[0] == "ALICE-Codec Bitstream Info"                      # main.rs:183-193
    assert f"  File size:   {len(alc)} bytes" in lines
    assert "  Width:       8" in lines and "  Height:      8" in lines and "  Frames:      2" in lines
    assert "  Wavelet:     CDF 9/7" in lines
    assert f"  Payload:     {len(alc) - 3138} bytes" in lines
    assert "  Raw size:    384 bytes (uncompressed RGB)" in lines
    assert f"  Ratio:       {(len(alc) - 3138) / 384 * 100:.1f}%" in lines
    # a two-chunk stream
    (tmp_path / "s.alc").write_bytes(alc + O.encode(rgb, 8, 8, 2, 50, 0))
    assert "  Chunks:      2 (multi-chunk stream)" in run("info", str(tmp_path / "s.alc")).stdout
    # errors: exit code 1 and "error: ..." on stderr (main.rs:104-107)
    r = run("info", str(tmp_path / "missing.alc"))
    assert r.returncode == 1 and r.stderr.startswith("error: ")
    (tmp_path / "bad.alc").write_bytes(b"XLCC" + alc[4:])
    r = run("info", str(tmp_path / "bad.alc"))
    assert r.returncode == 1 and "InvalidBitstream" in r.stderr
    (tmp_path / "a.rgb").write_bytes(rgb.tobytes())
    r = run("encode", str(tmp_path / "a.rgb"), "-W", "8", "-H", "8", "-f", "2", "-w", "dct", "-o", str(tmp_path / "x.alc"))
    assert r.returncode == 1 and "unknown wavelet 'dct'; expected cdf53, cdf97, or haar" in r.stderr
    # a single blob with trailing junk is still one chunk (from_bytes ignores trailing bytes, pipeline.rs:303)
    (tmp_path / "junk.alc").write_bytes(alc + b"trailing bytes that are not a blob" * 100)
    r = run("info", str(tmp_path / "junk.alc"))
    assert r.returncode == 0 and "Chunks:" not in r.stdout and "  Width:       8" in r.stdout
    # the reference's --quality is a u8: out-of-range values are an error, not a Python traceback
    r = run("encode", str(tmp_path / "a.rgb"), "-W", "8", "-H", "8", "-f", "2", "-q", "300", "-o", str(tmp_path / "x.alc"))
    assert r.returncode == 1 and r.stderr.startswith("error: ") and "Traceback" not in r.stderr


@pytest.mark.gpu
def test_encode_decode_files_match_the_oracle(tmp_path):
    w, h, f = 48, 20, 12
    rgb = O.generate(O.G1, w, h, f)
    (tmp_path / "in.rgb").write_bytes(rgb.tobytes())
    r = run("encode", str(tmp_path / "in.rgb"), "-W", str(w), "-H", str(h), "-f", str(f), "-q", "80", "-w", "cdf97",
            "-o", str(tmp_path / "out.alc"))
    assert r.returncode == 0, r.stderr
    ref = O.encode(rgb, w, h, f, 80, 1)
    assert (tmp_path / "out.alc").read_bytes() == ref
    assert r.stderr.startswith(f"encoded {w}x{h}x{f} ({rgb.size} bytes) -> {len(ref)} bytes (")
    r = run("decode", str(tmp_path / "out.alc"), "-o", str(tmp_path / "back.rgb"))
    assert r.returncode == 0, r.stderr
    assert np.array_equal(np.fromfile(tmp_path / "back.rgb", dtype=np.uint8), O.decode(ref))
    assert f"decoded {w}x{h}x{f} -> {rgb.size} bytes (raw RGB)" in r.stderr
    # multi-chunk stream: 12 frames in chunks of 5 -> 5 + 5 + 2
    r = run("encode", str(tmp_path / "in.rgb"), "-W", str(w), "-H", str(h), "-f", str(f), "--chunk-frames", "5",
            "-o", str(tmp_path / "stream.alc"))
    assert r.returncode == 0, r.stderr
    per = w * h * 3
    refs = [O.encode(rgb[t0 * per:min(t0 + 5, f) * per], w, h, min(5, f - t0), 90, 0) for t0 in range(0, f, 5)]
    assert (tmp_path / "stream.alc").read_bytes() == b"".join(refs)
    r = run("decode", str(tmp_path / "stream.alc"), "-o", str(tmp_path / "stream.rgb"))
    assert r.returncode == 0, r.stderr
    assert np.array_equal(np.fromfile(tmp_path / "stream.rgb", dtype=np.uint8), np.concatenate([O.decode(b) for b in refs]))
    # wrong size: InvalidBufferSize, exit code 1
    r = run("encode", str(tmp_path / "in.rgb"), "-W", str(w), "-H", str(h), "-f", str(f + 1), "-o", str(tmp_path / "z.alc"))
    assert r.returncode == 1 and "InvalidBufferSize" in r.stderr
