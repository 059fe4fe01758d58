"""CPU tier: the CUDA kernel sources compiled against the SIMT emulator (tests/emul/cuda_emul.h, test
infrastructure) and driven through the same C ABI, compared bit for bit with the oracle.  This debugs kernel
*logic* where no GPU exists; the -m gpu tier (test_gpu_parity.py) is the parity proof on real hardware."""
import importlib.util
import os

import numpy as np
import pytest

import oracle as O
import parity
from ref_vectors import SURVEY_KATS

pkg = parity.pkg


@pytest.fixture(scope="module")
def api():
    spec = importlib.util.spec_from_file_location("b", os.path.join(os.path.dirname(pkg.__file__), "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    return pkg.Api(b.build_emul())


@pytest.mark.parametrize("row", SURVEY_KATS, ids=lambda r: f"G{r[0]}-{r[1]}x{r[2]}x{r[3]}-q{r[4]}-w{r[5]}")
def test_survey_kats(api, row):
    parity.check_kat(api, row)


@pytest.mark.parametrize("shape", [(1, 1, 1), (2, 2, 2), (3, 5, 1), (5, 3, 3), (7, 2, 4), (2, 9, 5), (66, 6, 2),
                                   (130, 4, 2), (6, 70, 2), (12, 6, 64), (256, 6, 2), (380, 10, 3), (260, 4, 64),
                                   (80, 2, 64), (128, 6, 64), (112, 40, 64)])   # the last three: the fused front-end kernel
@pytest.mark.parametrize("wavelet", [0, 1, 2])
def test_odd_and_edge_shapes(api, shape, wavelet):
    w, h, f = shape
    parity.check_encode_decode(api, O.G1, w, h, f, 80, wavelet)


@pytest.mark.parametrize("kind,quality", [(O.G2, 100), (O.G2, 0), (O.G0, 100), (O.G1, 50)])
def test_noise_and_extreme_quality(api, kind, quality):
    for wavelet in (0, 1, 2):
        parity.check_encode_decode(api, kind, 24, 10, 6, quality, wavelet)


def test_random_shapes(api):
    """seeded random (shape, wavelet, quality, input kind) cases: strips with and without interior lanes, 1-frame
    to 66-frame chunks (the 64-deep chunks run the compile-time-depth temporal kernels), every stage compared"""
    rng = np.random.default_rng(20261018)
    done = 0
    while done < 24:
        w = int(rng.choice([rng.integers(1, 40), rng.integers(40, 300), rng.integers(240, 520)]))
        h = int(rng.integers(1, 48))
        f = int(rng.choice([1, 2, 3, 5, 7, 62, 63, 64, 65, 66]))
        if w * h * f > 300000:
            continue
        parity.check_encode_decode(api, int(rng.choice([O.G0, O.G1, O.G2])), w, h, f,
                                   int(rng.choice([0, 30, 75, 80, 90, 100])), int(rng.integers(0, 3)))
        done += 1


def test_stage_apis(api):
    rng = np.random.default_rng(1)
    parity.check_wavelet_api(api, rng, [2, 3, 8, 9, 31], [(4, 4), (5, 3), (16, 9)], [(4, 4, 4), (5, 3, 2), (8, 6, 3)])
    parity.check_wavelet_extremes(api)
    parity.check_quant_api(api, rng, n=3000)
    parity.check_colour(api, rng, n=1000)
    parity.check_rdo(api, rng)


def test_rdo_exact_sum_and_octants(api):
    rng = np.random.default_rng(7)
    parity.check_rdo_exact_variance(api, rng, sizes=(1, 2, 31, 1024, 1025, 5000))
    parity.check_rdo_octants(api, rng, shapes=((8, 6, 4), (9, 7, 5), (2, 2, 2), (32, 18, 8)))


def test_psnr_device_kernel(api):
    """the emulator's device pointers are host pointers, so the squared-difference kernel can be driven from numpy"""
    rng = np.random.default_rng(3)
    for n in (1, 15, 16, 4097, 50001):
        a = rng.integers(0, 256, n + 16, dtype=np.uint8)
        b = rng.integers(0, 256, n + 16, dtype=np.uint8)
        for off in (0, 1):
            x, y = a[off:off + n], b[off:off + n]
            assert api.psnr_device(x.ctypes.data, y.ctypes.data, n) == O.psnr(x, y) == api.psnr(x, y)
        assert api.psnr_device(a.ctypes.data, a.ctypes.data, n) == float("inf")
    assert api.psnr_device(0, 0, 0) == float("inf")


def test_rans_api(api):
    parity.check_rans_api(api, np.random.default_rng(2), n=3000)


def test_rans_decode_split16_layout(api, monkeypatch):
    """the 16-bit split decode tables (33 KB per stream, six streams per SM; off by default until measured) decode
    exactly like the 8-byte entries: rANS API cases, malformed tables included, and whole chunks"""
    monkeypatch.setenv("ALICE_RANS_DEC_SPLIT16", "1")
    parity.check_rans_api(api, np.random.default_rng(2), n=3000)
    for wavelet in (0, 1, 2):
        parity.check_encode_decode(api, O.G1, 260, 4, 64, 80, wavelet)
        parity.check_encode_decode(api, O.G2, 24, 10, 6, 100, wavelet)
    parity.check_decode_foreign_headers(api, np.random.default_rng(3))


def test_rans_interleaved(api):
    parity.check_rans_interleaved(api, np.random.default_rng(5), sizes=(0, 1, 2, 3, 4, 5, 7, 1024, 4099))


def test_errors_and_abi(api):
    parity.check_errors(api)
    parity.check_reference_abi(api)
    parity.check_decode_foreign_headers(api, np.random.default_rng(3))


def test_shared_workspace_batch(api):
    parity.check_shared_workspace_batch(api, shapes=((20, 12, 6), (21, 13, 5), (96, 4, 64)), n=3)


def test_payload_arena_tight_and_overflowing(api):
    parity.check_payload_arena(api)


def test_wavelet_api_fast_path(api):
    parity.check_wavelet_fast_path(api)


def test_lossless_set(api):
    parity.check_lossless_set(api)


def test_batch_submit_collect(api):
    parity.check_submit_collect(api)


def test_shifted_in_place_batch(api):
    parity.check_shifted_in_place(api)


def test_stream_device_batch(api):
    parity.check_stream_device(api)
