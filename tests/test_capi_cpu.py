"""CPU tier: the C-ABI library loads and exports every symbol include/alice_codec.h declares, the host-side
logic that needs no device behaves like the reference, and compute entry points fail loudly without a GPU."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle as O
from __graft_entry__ import load_package

pkg = load_package()


@pytest.fixture(scope="module")
def api():
    import importlib.util
    spec = importlib.util.spec_from_file_location("b", os.path.join(os.path.dirname(pkg.__file__), "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    b.build()
    return pkg.Api()


def test_exports_every_declared_symbol(api):
    declared = pkg._capi.declared_symbols()
    assert len(declared) >= 60
    assert sorted(pkg._capi.SIGNATURES) == declared, "ctypes table and header disagree"
    for name in declared:
        assert hasattr(api.lib, name), f"{name} declared in include/alice_codec.h but not exported"


def test_reference_abi_symbol_list(api):
    # the 20 symbols of src/ffi.rs:16-315 (bindings/ue5/AliceCodec.h:14-68)
    ref = ["wavelet1d_haar", "wavelet1d_cdf53", "wavelet1d_cdf97", "wavelet1d_destroy", "wavelet1d_forward",
           "wavelet1d_inverse", "encoder_create", "encoder_destroy", "encode", "decode", "chunk_destroy",
           "chunk_to_bytes", "chunk_from_bytes", "chunk_width", "chunk_height", "chunk_frames", "psnr", "data_free",
           "string_free", "version"]
    assert len(ref) == 20
    for n in ref:
        assert hasattr(api.lib, "alice_codec_" + n)


def test_host_only_entry_points(api):
    assert api.version() == "0.1.2"
    good = O.encode(O.generate(O.G0, 4, 4, 2), 4, 4, 2, 90, 0)
    ck = pkg.EncodedChunk.from_bytes(good, api=api)                       # pipeline.rs:235-313, no device needed
    assert (ck.width, ck.height, ck.frames, ck.wavelet) == (4, 4, 2, "cdf53")
    assert ck.to_bytes() == good
    assert ck.compressed_size == len(good) - 3138
    hdr = ck.channel_header(0)
    assert hdr["num_symbols"] == 32 and hdr["quant_step"] == 8 and hdr["quant_dead_zone"] == 8
    with pytest.raises(pkg.CodecError) as e:
        pkg.EncodedChunk.from_bytes(good[:3000], api=api)
    assert e.value.kind == "InvalidBitstream"
    for q in (0, 50, 75, 80, 90, 100, 255):
        assert api.rdo_bpp_from_quality(q) == O.rdo_bpp_from_quality(q)
    # empty chunk and argument validation happen before any device work (pipeline.rs:384-427)
    enc = pkg.FrameEncoder(90, "cdf97", api=api)
    assert enc.encode(np.zeros(0, np.uint8), 0, 5, 5).to_bytes() == O.encode(np.zeros(0, np.uint8), 0, 5, 5, 90, 1)
    with pytest.raises(pkg.CodecError) as e:
        enc.encode(np.zeros(5, np.uint8), 2, 2, 2)
    assert e.value.kind == "InvalidBufferSize"
    with pytest.raises(ValueError):
        pkg.FrameEncoder(90, "dct", api=api)                              # python.rs:385-389


def test_no_cpu_fallback(api):
    """Without a CUDA device every compute entry point must fail loudly (ALICE_ERR_CUDA), never compute."""
    if api.device_count() > 0:
        pytest.skip("a CUDA device is present")
    rgb = O.generate(O.G0, 4, 4, 2)
    with pytest.raises(pkg.CodecError) as e:
        pkg.FrameEncoder(90, "cdf53", api=api).encode(rgb, 4, 4, 2)
    assert e.value.kind == "CudaError"
    with pytest.raises(pkg.CodecError) as e:
        pkg.FrameDecoder(api=api).decode(pkg.EncodedChunk.from_bytes(O.encode(rgb, 4, 4, 2, 90, 0), api=api))
    assert e.value.kind == "CudaError"
    for fn in (lambda: api.to_symbols([1, 2]), lambda: api.wavelet3d(0, np.zeros(8, np.int32), 2, 2, 2),
               lambda: api.build_histogram([1]), lambda: api.rans_encode([1], [1, 1]),
               lambda: api.quantize_buffer(8, 8, [1]), lambda: api.rgb_to_ycocg_r([1, 2, 3])):
        with pytest.raises(pkg.CodecError) as e:
            fn()
        assert e.value.kind == "CudaError"
    assert pkg.ReferenceAbi(api).encode_to_bytes(rgb, 4, 4, 2) is None      # reference ABI: null


def test_product_does_not_reference_the_oracle():
    """The product tree must not include, link or import anything under oracle/."""
    root = os.path.dirname(pkg.__file__)
    for dp, _, files in os.walk(root):
        if os.path.basename(dp) in ("lib", "obj", "__pycache__"):
            continue
        for fn in files:
            if fn.endswith((".cu", ".cuh", ".h", ".py")):
                text = open(os.path.join(dp, fn), errors="ignore").read()
                assert "alice_oracle" not in text and "import oracle" not in text and "from oracle" not in text, fn
    deps = os.popen(f"ldd {pkg._capi.PRODUCT_LIB}").read()
    assert "oracle" not in deps
