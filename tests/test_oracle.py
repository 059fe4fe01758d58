"""CPU tests that pin the oracle (oracle/alice_oracle.c) before anything trusts it:
 * every exact-value assertion the reference's own tests hold for the hot path, and
 * the SURVEY.md Appendix C known-answer vectors (sha256 of .alc bytes / decoded RGB).
"""
import hashlib

import numpy as np
import pytest

import oracle as O
from ref_vectors import (FREQ_4BIN, KAT_4x4x2_HEADER_HEX, KAT_4x4x2_STREAMS_HEX, KAT_4x4x2_Y_COEFFS,
                         KAT_4x4x2_Y_SYMBOLS, LOSSLESS_EXACT_1D, LOSSLESS_EXACT_2D, SURVEY_KATS,
                         W1D_ROUNDTRIP_53, W1D_VECTORS)


# ---- quant.rs
def test_quantizer_doc_example():  # quant.rs:52-54
    assert O.quantize(8, 8, 20) == 2
    assert O.dequantize(8, 2) == 16


def test_dead_zone():  # quant.rs:729-737, 866-875
    for v in range(-15, 16):
        assert O.quantize(16, 16, v) == 0
        assert O.fast_quantize_buffer(16, 16, [v])[0] == 0


def test_symbol_ordering_and_roundtrip():  # quant.rs:741-765
    assert list(O.to_symbols([0, 1, -1, 2, -2, 3, -3])) == [0, 1, 2, 3, 4, 5, 6]
    orig = [-5, -2, -1, 0, 1, 2, 5]
    assert list(O.from_symbols(O.to_symbols(orig))) == orig


def test_symbol_wraps_mod_256():  # quant.rs:555-560 (`as u8` truncation)
    assert list(O.to_symbols([128, 129, -128, -129, 1000])) == [255, 1, 0, 2, (2 * 1000 - 1) & 255]


def test_histogram():  # quant.rs:804-813
    h = O.build_histogram([0, 0, 1, 1, 1, 2, 5, 5])
    assert (h[0], h[1], h[2], h[3], h[5]) == (2, 3, 1, 0, 2)
    assert h.sum() == 8


def test_dequantize_buffer_exact():  # quant.rs:1088-1098
    assert list(O.dequantize_buffer(8, [0, 1, -1, 5, -5])) == [0, 8, -8, 40, -40]


def test_fast_quantizer_matches_regular():  # quant.rs:848-864, 1145-1150
    vals = np.arange(-10000, 10001, dtype=np.int32)
    for step in list(range(1, 129)):
        assert np.array_equal(O.quantize_buffer(step, step, vals), O.fast_quantize_buffer(step, step, vals))


def test_fast_quantizer_invalid_step():  # quant.rs:1116-1122
    for s in (0, -5):
        with pytest.raises(O.OracleError) as e:
            O.fast_quantize_buffer(s, s, [1])
        assert e.value.code == O.ERR_QUANT_STEP


def test_rdo_basic():  # quant.rs:767-801
    bpp50 = O.rdo_bpp_from_quality(50)
    coeffs = np.arange(-100, 101, dtype=np.int32)
    s_lll, dz = O.rdo_compute_quantizer(bpp50, coeffs, 0)
    s_hhh, _ = O.rdo_compute_quantizer(bpp50, coeffs, 7)
    assert s_lll > 0 and s_hhh >= s_lll and s_hhh == 8 * s_lll and dz == s_lll + s_lll // 2
    assert O.rdo_bpp_from_quality(10) < O.rdo_bpp_from_quality(90)
    assert O.rdo_estimate_variance([]) == 1.0


# ---- color.rs
def test_color_roundtrip_lattice():  # color.rs:429-461
    g = np.arange(0, 256, 17, dtype=np.uint8)
    rgb = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    extra = np.array([[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255], [1, 2, 3]], np.uint8)
    rgb = np.concatenate([rgb, extra]).ravel()
    y, co, cg = O.rgb_bytes_to_ycocg_r(rgb)
    assert np.array_equal(O.ycocg_r_to_rgb_bytes(y, co, cg), rgb)


def test_color_known_values():  # color.rs:557-606
    y, co, cg = O.rgb_bytes_to_ycocg_r(np.array([255, 0, 0, 77, 77, 77], np.uint8))
    assert co[0] == 255
    assert (y[1], co[1], cg[1]) == (77, 0, 0)


# ---- wavelet.rs / lossless.rs
@pytest.mark.parametrize("wt,inp,fwd", W1D_VECTORS)
def test_wavelet1d_vectors(wt, inp, fwd):
    assert list(O.wavelet1d_forward(wt, inp)) == fwd


def test_wavelet1d_regression_roundtrip():  # proptest-regressions/wavelet.txt:7 (inexact by design)
    inp, back = W1D_ROUNDTRIP_53
    assert list(O.wavelet1d_inverse(0, O.wavelet1d_forward(0, inp))) == back


@pytest.mark.parametrize("wt,tol", [(2, 1), (0, 1), (1, 2)])
def test_wavelet1d_roundtrip_tolerance(wt, tol):  # wavelet.rs:492-531
    sig = [10, 20, 30, 40, 50, 60, 70, 80]
    back = O.wavelet1d_inverse(wt, O.wavelet1d_forward(wt, sig))
    assert np.abs(back - np.array(sig)).max() <= tol


@pytest.mark.parametrize("sig", LOSSLESS_EXACT_1D)
def test_lossless_exact_1d(sig):  # lossless.rs:110-159,178-185
    assert list(O.wavelet1d_inverse(0, O.wavelet1d_forward(0, sig))) == list(sig)


@pytest.mark.parametrize("img,w,h", LOSSLESS_EXACT_2D)
def test_lossless_exact_2d(img, w, h):  # lossless.rs:149-159
    assert list(O.wavelet2d_inverse(0, O.wavelet2d_forward(0, img, w, h), w, h)) == list(img)


def test_wavelet3d_roundtrip_tolerance():  # wavelet.rs:551-563
    vol = np.array([(i * 3) % 256 for i in range(4 * 4 * 2)], np.int32)
    back = O.wavelet3d_inverse(0, O.wavelet3d_forward(0, vol, 4, 4, 2), 4, 4, 2)
    assert np.abs(back - vol).max() <= 3


# ---- rans.rs
def test_uniform_table():  # rans.rs:719-735, 941-951
    t = O.freq_table_uniform(256)
    assert set(t.freq_np()) == {16}
    t2 = O.freq_table_uniform(2)
    assert t2.cum[0] == 0 and t2.freq[0] + t2.freq[1] == 4096


def test_histogram_normalization():  # rans.rs:819-830
    hist, freq, cum = FREQ_4BIN
    t = O.freq_table_from_histogram(hist)
    assert list(t.freq_np()[:4]) == freq and list(t.cum_np()[:4]) == cum


def test_malformed_table_single_bin():  # SURVEY A.10 worked example (rans.rs:117-132)
    h = np.zeros(256, np.uint32)
    h[100] = 1000
    t = O.freq_table_from_histogram(h)
    assert t.freq[100] == 4096 and t.cum[100] == 100 and t.freq[255] == 65282


@pytest.mark.parametrize("syms", [[42, 100, 200, 50, 128], [0], [42] * 500, list(range(100)), []])
def test_rans_roundtrip_uniform(syms):  # rans.rs:738-751, 853-880, 925-935
    t = O.freq_table_uniform(256)
    enc = O.rans_encode(syms, t)
    assert list(O.rans_decode(enc, len(syms), t)) == list(syms)
    if not syms:
        assert enc == bytes([0x00, 0x80, 0x00, 0x00])


def test_rans_roundtrip_skewed():  # rans.rs:754-787
    h = np.ones(256, np.uint32)
    h[0], h[1], h[2] = 1000, 500, 100
    t = O.freq_table_from_histogram(h)
    syms = [0 if i % 10 <= 6 else (1 if i % 10 <= 8 else 2) for i in range(1000)]
    enc = O.rans_encode(syms, t)
    assert len(enc) < len(syms)
    assert list(O.rans_decode(enc, len(syms), t)) == syms


# ---- pipeline.rs
def test_quality_to_step():  # pipeline.rs:456-457
    assert [O.quality_to_step(q) for q in (90, 80, 75, 50, 100, 0, 255)] == [8, 14, 17, 33, 1, 64, 1]


@pytest.mark.parametrize("w,h,f,q,wt,floor", [(4, 4, 2, 90, 0, 15.0), (8, 8, 2, 100, 2, 5.0), (3, 4, 2, 90, 0, 10.0),
                                               (4, 5, 2, 90, 0, 10.0), (3, 5, 1, 90, 0, 10.0)])
def test_pipeline_psnr_floors(w, h, f, q, wt, floor):  # pipeline.rs:686-829, 860-878
    rgb = O.generate(O.G0, w, h, f)
    dec = O.decode(O.encode(rgb, w, h, f, q, wt))
    assert dec.size == rgb.size
    assert O.psnr(rgb, dec) > floor


def test_pipeline_solid_color():  # pipeline.rs:696-709
    rgb = np.tile(np.array([128, 64, 200], np.uint8), 4 * 4 * 2)
    dec = O.decode(O.encode(rgb, 4, 4, 2, 95, 0))
    assert O.psnr(rgb, dec) > 25.0


def test_pipeline_errors():  # pipeline.rs:786-797, 391-412
    with pytest.raises(O.OracleError) as e:
        O.encode(np.zeros(10, np.uint8), 4, 4, 2, 50, 0)
    assert e.value.code == O.ERR_BUFFER_SIZE
    with pytest.raises(O.OracleError) as e:
        O.encode(np.zeros(0, np.uint8), 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 50, 0)
    assert e.value.code == O.ERR_OVERFLOW
    empty = O.encode(np.zeros(0, np.uint8), 0, 4, 2, 50, 0)
    assert len(empty) == 3138 and O.decode(empty).size == 0
    with pytest.raises(O.OracleError):
        O.decode(b"ALCX" + bytes(3200))
    with pytest.raises(O.OracleError):
        O.decode(bytes(100))


def test_pipeline_1x1x1():  # pipeline.rs:832-844
    dec = O.decode(O.encode(np.array([128, 200, 50], np.uint8), 1, 1, 1, 100, 0))
    assert dec.size == 3


# ---- SURVEY.md Appendix C
@pytest.mark.parametrize("kind,w,h,f,q,wt,alc_len,sha_alc,sha_rgb", SURVEY_KATS)
def test_survey_kats(kind, w, h, f, q, wt, alc_len, sha_alc, sha_rgb):
    rgb = O.generate(kind, w, h, f)
    alc = O.encode(rgb, w, h, f, q, wt)
    assert len(alc) == alc_len
    assert hashlib.sha256(alc).hexdigest() == sha_alc
    assert hashlib.sha256(O.decode(alc).tobytes()).hexdigest() == sha_rgb


def test_survey_kat_expanded():
    rgb = O.generate(O.G0, 4, 4, 2)
    alc, coeffs, syms = O.encode(rgb, 4, 4, 2, 90, 0, stages=True)
    assert alc[:18].hex() == KAT_4x4x2_HEADER_HEX
    assert alc[18:34].hex() == "0c000000080000000800000020000000"
    assert list(coeffs[0]) == KAT_4x4x2_Y_COEFFS
    assert list(syms[0]) == KAT_4x4x2_Y_SYMBOLS
    assert alc[3138:].hex() == "".join(KAT_4x4x2_STREAMS_HEX)


def test_generator_g1_prefix():
    assert list(O.generate(O.G1, 256, 128, 64)[:12]) == [63, 90, 194, 66, 94, 186, 66, 92, 184, 67, 96, 184]
    assert hashlib.sha256(O.generate(O.G1, 64, 32, 8).tobytes()).hexdigest().startswith("a7977982a603b092a95cce271f9b0c89")
