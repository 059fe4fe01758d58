"""GPU tier (-m gpu): the parity tests proper.  The in-tree CUDA library (sm_100a) is driven through its C ABI and
compared bit for bit with the CPU oracle on the same seeded inputs; BASELINE.json's full sizes are checked
against committed oracle digests (tests/golden/fullsize.json) and through size-independent properties."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle as O
import parity
from ref_vectors import SURVEY_KATS

pkg = parity.pkg
pytestmark = pytest.mark.gpu
GOLDEN_PATH = os.path.join(os.path.dirname(__file__), "golden", "fullsize.json")


def sha(b):
    return hashlib.sha256(b).hexdigest()


@pytest.fixture(scope="module")
def api():
    a = pkg.Api()                      # the product library; raises if it is missing (no fallback)
    assert a.path.endswith("alice-codec_b200/lib/libalice_codec.so")
    assert a.device_count() >= 1, "no CUDA device visible to libalice_codec"
    a.set_device(0)
    return a


@pytest.mark.parametrize("row", SURVEY_KATS, ids=lambda r: f"G{r[0]}-{r[1]}x{r[2]}x{r[3]}-q{r[4]}-w{r[5]}")
def test_survey_kats(api, row):
    parity.check_kat(api, row)


SHAPES = [(1, 1, 1), (2, 2, 2), (3, 5, 1), (5, 3, 3), (7, 2, 4), (2, 9, 5), (1, 40, 7), (40, 1, 7), (66, 6, 2),
          (130, 4, 2), (6, 70, 2), (12, 6, 64), (121, 67, 9), (124, 62, 64), (250, 30, 3), (64, 64, 65), (31, 33, 128), (256, 6, 2), (380, 10, 3), (260, 4, 64), (512, 270, 3),
          (80, 2, 64), (128, 6, 64), (112, 40, 64), (496, 64, 64), (1920, 24, 64)]   # last five: fused front-end kernel (w % 16 == 0, 64 frames)


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("wavelet", [0, 1, 2])
def test_odd_and_edge_shapes(api, shape, wavelet):
    w, h, f = shape
    parity.check_encode_decode(api, O.G1, w, h, f, 80, wavelet)


@pytest.mark.parametrize("kind", [O.G0, O.G1, O.G2])
@pytest.mark.parametrize("quality", [0, 50, 75, 80, 90, 100])
def test_every_quality_and_input_kind(api, kind, quality):
    for wavelet in (0, 1, 2):
        parity.check_encode_decode(api, kind, 96, 40, 16, quality, wavelet)


@pytest.mark.parametrize("cfg", [(90, 0), (80, 1), (75, 2)], ids=["cdf53-q90", "cdf97-q80", "haar-q75"])
@pytest.mark.parametrize("shape", [(256, 128, 64), (640, 360, 16), (1920, 1080, 2)], ids=lambda s: "x".join(map(str, s)))
def test_baseline_configs_medium(api, cfg, shape):
    """BASELINE.json configs 1-3 at sizes the oracle finishes in seconds (full stage-by-stage comparison)."""
    q, wv = cfg
    w, h, f = shape
    parity.check_encode_decode(api, O.G1, w, h, f, q, wv)
    parity.check_encode_decode(api, O.G2, w // 2, h // 2, f, q, wv)


def _fullsize(api, name, want_coeffs):
    g = json.load(open(GOLDEN_PATH))[name]
    rgb = O.generate(g["kind"], g["w"], g["h"], g["f"], g["seed"])
    assert sha(rgb.tobytes()) == g["sha256_rgb_in"]
    enc = pkg.FrameEncoder(g["quality"], g["wavelet"], api=api)
    chunk, coeffs, syms = enc.encode_stages(rgb, g["w"], g["h"], g["f"], want_coeffs=want_coeffs)
    del rgb
    n = syms.shape[1]
    for c in range(3):
        if want_coeffs:
            assert sha(coeffs[c].tobytes()) == g["sha256_coeffs"][c], f"coefficients differ (channel {c})"
        assert sha(syms[c].tobytes()) == g["sha256_symbols"][c], f"symbols differ (channel {c})"
        hdr = chunk.channel_header(c)
        assert int(hdr["histogram"].sum(dtype=np.uint64)) == n == hdr["num_symbols"]      # property: counts add up
        assert np.array_equal(hdr["histogram"], np.bincount(syms[c], minlength=256))      # property: hist == symbols
        assert hdr["compressed_len"] == g["stream_lens"][c]
    del coeffs
    alc = chunk.to_bytes()
    assert len(alc) == g["alc_len"]
    assert sha(alc[:3138]) == g["sha256_header"]
    assert sha(alc) == g["sha256_alc"], ".alc differs from the oracle digest"
    out, dsyms = pkg.FrameDecoder(api=api).decode_stages(pkg.EncodedChunk.from_bytes(alc, api=api))
    assert sha(out.tobytes()) == g["sha256_decoded"], "decoded RGB differs from the oracle digest"
    # property: a channel whose used symbols all sit inside [0,4096) of the table must survive the rANS round trip
    for c in range(3):
        cum, freq, _ = api.freq_table_from_histogram(chunk.channel_header(c)["histogram"])
        used = np.nonzero(chunk.channel_header(c)["histogram"])[0]
        if np.all(cum[used].astype(np.int64) + freq[used] <= 4096):
            assert np.array_equal(dsyms[c], syms[c]), f"well-formed stream {c} failed its round trip"


def test_fullsize_config1_cdf53_q90_1080p64(api):
    _fullsize(api, "cfg1_cdf53_q90_1080p64", want_coeffs=True)


def test_fullsize_config2_cdf97_q80_1080p64(api):
    _fullsize(api, "cfg2_cdf97_q80_1080p64", want_coeffs=True)


def test_fullsize_config3_haar_q75_4k64(api):
    _fullsize(api, "cfg3_haar_q75_4k64", want_coeffs=False)


def test_fullsize_odd_and_noise(api):
    _fullsize(api, "odd_cdf97_q80_1919x1079x63", want_coeffs=False)
    _fullsize(api, "noise_cdf53_q90_1080p8", want_coeffs=True)


def test_fullsize_config5_all_eight_4k_chunks_through_the_batch_api(api):
    """BASELINE config 5 (3840x2160x512 = 8 chunks of 64 frames, CDF 9/7 q=80, chunk c = G1 seed + c): all eight
    chunks in flight through the batch API, .alc and decoded RGB against the oracle digests."""
    import ctypes as C
    import torch
    g = json.load(open(GOLDEN_PATH))
    names = [f"cfg5_cdf97_q80_4k64_chunk{c}" for c in range(8)]
    w, h, f = 3840, 2160, 64
    st = torch.cuda.current_stream()
    d_in = [torch.empty(w * h * f * 3, dtype=torch.uint8, device="cuda") for _ in names]
    for t, nm in zip(d_in, names):
        api._chk(api.lib.alice_codec_synth_rgb_device(g[nm]["kind"], g[nm]["seed"], w, h, f, C.c_void_p(t.data_ptr()),
                                                      C.c_void_p(st.cuda_stream)))
    batch = pkg.ChunkBatch(80, "cdf97", w, h, f, len(names), stream=st.cuda_stream, api=api)
    batch.encode_device([t.data_ptr() for t in d_in])
    for i, nm in enumerate(names):
        if i in (0, 7):
            assert sha(d_in[i].cpu().numpy().tobytes()) == g[nm]["sha256_rgb_in"]
        alc = batch.get_chunk(i).to_bytes()
        assert len(alc) == g[nm]["alc_len"]
        assert sha(alc) == g[nm]["sha256_alc"], f"{nm}: .alc differs from the oracle digest"
    batch.decode_device([t.data_ptr() for t in d_in])        # in place: decode does not read the RGB input
    torch.cuda.synchronize()
    for i, nm in enumerate(names):
        assert sha(d_in[i].cpu().numpy().tobytes()) == g[nm]["sha256_decoded"], f"{nm}: decoded RGB differs"
    batch.close()


def test_stage_apis(api):
    rng = np.random.default_rng(1)
    parity.check_wavelet_api(api, rng, [2, 3, 8, 9, 31, 1000, 1921], [(4, 4), (5, 3), (16, 9), (321, 65)],
                             [(4, 4, 4), (5, 3, 2), (8, 6, 3), (64, 36, 64), (33, 17, 9)])
    parity.check_wavelet_extremes(api)
    parity.check_quant_api(api, rng, n=200000)
    parity.check_colour(api, rng, n=100000)
    parity.check_rdo(api, rng)


def test_rdo_exact_sum_and_octants(api):
    """AnalyticalRDO statistics with the sequential f64 sum reproduced bit for bit by the parallel
    binade scan (csrc/k_rdo.cu), compute_all_quantizers on the octants of a volume and the per-octant FastQuantizer."""
    rng = np.random.default_rng(7)
    parity.check_rdo_exact_variance(api, rng, sizes=(1, 2, 31, 1024, 1025, 5000, 40000, 1 << 20))
    parity.check_rdo_octants(api, rng)
    # a real forward-transformed volume: Y plane of G1, CDF 9/7
    w, h, d = 64, 36, 64
    y = O.rgb_bytes_to_ycocg_r(O.generate(O.G1, w, h, d))[0].astype(np.int32)
    vol = O.wavelet3d_forward(1, y, w, h, d)
    bpp = O.rdo_bpp_from_quality(80)
    want = [O.rdo_compute_quantizer(bpp, o, sb) for sb, o in enumerate(parity._octants(vol, w, h, d))]
    assert api.rdo_compute_all_quantizers(bpp, vol, w, h, d) == want


def test_psnr_device_equals_host(api):
    """alice_codec_psnr (ffi.rs:270) computed on the device for decode validation."""
    import torch
    rng = np.random.default_rng(3)
    for n in (1, 15, 16, 4097, 3 * 640 * 360 * 2):
        a = rng.integers(0, 256, n, dtype=np.uint8)
        b = np.clip(a.astype(np.int16) + rng.integers(-9, 10, n), 0, 255).astype(np.uint8)
        ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        got = api.psnr_device(ta.data_ptr(), tb.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
        assert got == api.psnr(a, b) == O.psnr(a, b), n
        assert api.psnr_device(ta.data_ptr(), ta.data_ptr(), n) == float("inf")
        # misaligned views take the byte path
        if n > 20:
            got = api.psnr_device(ta.data_ptr() + 1, tb.data_ptr() + 3, n - 3)
            assert got == O.psnr(a[1:n - 2], b[3:n])


def test_python_module_surface():
    """the names, defaults and error texts of the reference's Python module (src/python.rs:274-609) on top
    of the C ABI: FrameEncoder / FrameDecoder / EncodedChunk, rgb_to_ycocg_r_numpy, ycocg_r_to_rgb_numpy, version."""
    assert pkg.version() == "0.1.2"                                             # python.rs:274, Cargo.toml:3
    rgb = O.generate(O.G1, 16, 8, 2)
    y, co, cg = pkg.rgb_to_ycocg_r_numpy(rgb)
    ry, rco, rcg = O.rgb_bytes_to_ycocg_r(rgb)
    assert y.dtype == np.int16 and np.array_equal(y, ry) and np.array_equal(co, rco) and np.array_equal(cg, rcg)
    assert np.array_equal(pkg.ycocg_r_to_rgb_numpy(y, co, cg), rgb)
    with pytest.raises(ValueError, match="multiple of 3"):                       # python.rs:508-512
        pkg.rgb_to_ycocg_r_numpy(rgb[:-1])
    with pytest.raises(ValueError, match="same length"):                         # python.rs:551-555
        pkg.ycocg_r_to_rgb_numpy(y, co[:-1], cg)
    with pytest.raises(ValueError, match="C-contiguous"):                        # python.rs:505
        pkg.rgb_to_ycocg_r_numpy(np.zeros(12, np.uint8)[::2])
    with pytest.raises(ValueError, match="unknown wavelet type 'dwt'"):          # python.rs:385-389
        pkg.FrameEncoder(wavelet="dwt")
    enc = pkg.FrameEncoder()                                                     # quality=90, wavelet="cdf53" (python.rs:378)
    chunk = enc.encode(rgb, 16, 8, 2)
    assert chunk.to_bytes() == O.encode(rgb, 16, 8, 2, 90, O.CDF53)
    assert (chunk.width, chunk.height, chunk.frames, chunk.wavelet) == (16, 8, 2, "cdf53")
    assert repr(chunk) == f"EncodedChunk(16x8x2, {chunk.compressed_size} bytes, cdf53)"   # python.rs:343-355
    back = pkg.EncodedChunk.from_bytes(chunk.to_bytes())
    assert np.array_equal(pkg.FrameDecoder().decode(back), O.decode(chunk.to_bytes()))
    with pytest.raises(ValueError):                                              # CodecError -> ValueError (python.rs:61-63)
        enc.encode(rgb[:-3], 16, 8, 2)


def test_lossless_config4_transform(api):
    """BASELINE config 4: LosslessEncoder::transform_2d / inverse_2d (lossless.rs:45-54) == Wavelet2D::cdf53 per
    frame of the Y/Co/Cg planes; symbols (step 1, wrapping), histogram, table, rANS round trip per stream."""
    w, h, f = 480, 270, 4
    planes = O.rgb_bytes_to_ycocg_r(O.generate(O.G1, w, h, f))
    for p in planes:
        for t in range(f):
            img = p[t * w * h:(t + 1) * w * h].astype(np.int32)
            fw = api.wavelet2d(0, img, w, h)
            assert np.array_equal(fw, O.wavelet2d_forward(0, img, w, h))
            assert np.array_equal(api.wavelet2d(0, fw, w, h, inverse=True), O.wavelet2d_inverse(0, fw, w, h))
            if t == 0:
                sy = api.to_symbols(fw)
                assert np.array_equal(sy, O.to_symbols(fw))
                hist = api.build_histogram(sy)
                table = O.freq_table_from_histogram(hist)
                stream = api.rans_encode(sy, hist)
                assert stream == O.rans_encode(sy, table)
                assert np.array_equal(api.rans_decode(stream, sy.size, hist), O.rans_decode(stream, sy.size, table))


def test_rans_api(api):
    parity.check_rans_api(api, np.random.default_rng(2), n=300000)


def test_rans_interleaved(api):
    """the reference's 4-stream container (rans.rs:393-524) over the one-warp-per-stream kernels."""
    parity.check_rans_interleaved(api, np.random.default_rng(5), sizes=(0, 1, 2, 3, 4, 5, 7, 1024, 4099, 70001, 1 << 20))


def test_errors_and_abi(api):
    parity.check_errors(api)
    parity.check_reference_abi(api)
    parity.check_decode_foreign_headers(api, np.random.default_rng(3))


def test_batch_api_device_and_host(api):
    """The throughput path: many chunks of one shape in flight (alice_codec_batch_*), device and host variants."""
    import torch
    w, h, f, n = 128, 72, 16, 5
    for q, wv in [(90, "cdf53"), (80, "cdf97"), (75, "haar")]:
        batch = pkg.ChunkBatch(q, wv, w, h, f, n, stream=torch.cuda.current_stream().cuda_stream, api=api)
        rgbs = [O.generate(O.G1 if i != 3 else O.G2, w, h, f, O.SEED + i) for i in range(n)]
        d_in = [torch.from_numpy(r).cuda() for r in rgbs]
        d_out = [torch.zeros_like(t) for t in d_in]
        batch.encode_device([t.data_ptr() for t in d_in])
        refs = [O.encode(r, w, h, f, q, pkg.WAVELET_NAMES[wv]) for r in rgbs]
        for i in range(n):
            assert batch.get_chunk(i).to_bytes() == refs[i], f"chunk {i}"
        batch.decode_device([t.data_ptr() for t in d_out])
        torch.cuda.synchronize()
        for i in range(n):
            assert np.array_equal(d_out[i].cpu().numpy(), O.decode(refs[i])), f"decoded chunk {i}"
        # host variant (H2D / D2H inside the call)
        h_in = [torch.from_numpy(r).pin_memory() for r in rgbs]
        h_out = [torch.zeros(r.size, dtype=torch.uint8).pin_memory() for r in rgbs]
        chunks = batch.encode_host([t.data_ptr() for t in h_in])
        assert [c.to_bytes() for c in chunks] == refs
        batch.decode_host(chunks, [t.data_ptr() for t in h_out])
        for i in range(n):
            assert np.array_equal(h_out[i].numpy(), O.decode(refs[i]))
        # decode_host reuses the batch's histogram / payload buffers: what the last encode left resident is gone, and the
        # calls that would read it say so instead of returning another batch's data (include/alice_codec.h)
        with pytest.raises(pkg.CodecError) as ei:
            batch.get_chunk(0)
        assert ei.value.kind == "InvalidBufferSize"
        with pytest.raises(pkg.CodecError):
            batch.decode_device([t.data_ptr() for t in d_out])
        batch.encode_device([t.data_ptr() for t in d_in])          # ... until the next encode
        assert batch.get_chunk(n - 1).to_bytes() == refs[n - 1]
        assert len(batch.timings()) == 8 and batch.device_bytes() > 0
        batch.close()


def test_shared_workspace_batch(api):
    parity.check_shared_workspace_batch(api, shapes=((20, 12, 6), (21, 13, 5), (256, 64, 64), (96, 4, 64)), n=3)


def test_synth_on_device_matches_oracle_generators(api):
    import ctypes as C
    import torch
    for kind in (O.G0, O.G1, O.G2):
        for (w, h, f) in [(64, 32, 8), (33, 17, 5)]:
            t = torch.empty(w * h * f * 3, dtype=torch.uint8, device="cuda")
            api._chk(api.lib.alice_codec_synth_rgb_device(kind, O.SEED, w, h, f, C.c_void_p(t.data_ptr()), None))
            torch.cuda.synchronize()
            assert np.array_equal(t.cpu().numpy(), O.generate(kind, w, h, f))


def test_determinism_and_idempotence(api):
    w, h, f = 320, 180, 32
    rgb = O.generate(O.G1, w, h, f)
    enc = pkg.FrameEncoder(80, "cdf97", api=api)
    a = enc.encode(rgb, w, h, f).to_bytes()
    b = enc.encode(rgb, w, h, f).to_bytes()
    assert a == b
    ck = pkg.EncodedChunk.from_bytes(a, api=api)
    assert pkg.EncodedChunk.from_bytes(ck.to_bytes(), api=api).to_bytes() == a
    d1 = pkg.FrameDecoder(api=api).decode(ck)
    d2 = pkg.FrameDecoder(api=api).decode(ck)
    assert np.array_equal(d1, d2)


def test_payload_arena_tight_and_overflowing(api):
    parity.check_payload_arena(api)


def test_wavelet_api_fast_path(api):
    parity.check_wavelet_fast_path(api)


def test_lossless_set(api):
    parity.check_lossless_set(api)


def test_fullsize_config4_lossless_1080p64(api):
    """BASELINE config 4 at full size: every stage of the lossless frame-set pipeline against the oracle digests."""
    import ctypes as C
    import torch
    g = json.load(open(GOLDEN_PATH))["cfg4_lossless_1080p64"]
    w, h, f = g["w"], g["h"], g["f"]
    st = torch.cuda.current_stream()
    d_rgb = torch.empty(w * h * f * 3, dtype=torch.uint8, device="cuda")
    api._chk(api.lib.alice_codec_synth_rgb_device(g["kind"], g["seed"], w, h, f, C.c_void_p(d_rgb.data_ptr()), C.c_void_p(st.cuda_stream)))
    assert sha(d_rgb.cpu().numpy().tobytes()) == g["sha256_rgb_in"]
    ls = pkg.LosslessSet(w, h, f, stream=st.cuda_stream, api=api)
    ls.encode_device(d_rgb.data_ptr())
    ls.decode_device()
    for what, key in (("coeffs", "sha256_coeffs"), ("symbols", "sha256_symbols"), ("hist", "sha256_hist"),
                      ("decoded", "sha256_decoded_symbols"), ("inverse", "sha256_inverse")):
        got = ls.fetch(what)
        for c in range(3):
            assert sha(got[c].tobytes()) == g[key][c], f"{what} of channel {c} differs from the oracle digest"
    for c in range(3):
        s = ls.stream(c)
        assert len(s) == g["stream_lens"][c] and sha(s) == g["sha256_streams"][c], f"rANS stream {c}"
    ls.close()


def test_batch_submit_collect(api):
    parity.check_submit_collect(api)


def test_shifted_in_place_batch(api):
    parity.check_shifted_in_place(api)


def test_stream_device_batch(api):
    parity.check_stream_device(api)
