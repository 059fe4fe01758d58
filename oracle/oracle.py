"""ctypes binding of oracle/alice_oracle.c (CPU restatement of the reference; test-only).

Every wrapper cites the reference function it restates; see alice_oracle.c for file:line.
Parity status: pinned by the reference's exact-value tests and SURVEY.md Appendix C KATs;
the Rust reference itself cannot be built in this image (no rustc/cargo).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libalice_oracle.so")

OK, ERR_BUFFER_SIZE, ERR_DIMENSIONS, ERR_OVERFLOW, ERR_BITSTREAM, ERR_QUANT_STEP, ERR_PANIC = range(7)
ERR_NAMES = {
    1: "InvalidBufferSize", 2: "InvalidDimensions", 3: "DimensionOverflow",
    4: "InvalidBitstream", 5: "InvalidQuantStep", 6: "ReferencePanic",
}
CDF53, CDF97, HAAR = 0, 1, 2


class OracleError(ValueError):
    def __init__(self, code):
        super().__init__(ERR_NAMES.get(code, str(code)))
        self.code = code


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "alice_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _SO


class _W1D(C.Structure):
    _fields_ = [("n_steps", C.c_int), ("coeff", C.c_int32 * 4), ("predict", C.c_int * 4)]


class FreqTable(C.Structure):
    _fields_ = [("n_symbols", C.c_uint32), ("cum", C.c_uint16 * 256), ("freq", C.c_uint16 * 256),
                ("lut", C.c_uint8 * 4096)]

    def cum_np(self):
        return np.ctypeslib.as_array(self.cum).copy()

    def freq_np(self):
        return np.ctypeslib.as_array(self.freq).copy()

    def lut_np(self):
        return np.ctypeslib.as_array(self.lut).copy()


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        p8, p16, p32 = C.POINTER(C.c_uint8), C.POINTER(C.c_int16), C.POINTER(C.c_int32)
        sz = C.c_size_t
        L.alo_rgb_bytes_to_ycocg_r.argtypes = [p8, sz, p16, p16, p16, sz]
        L.alo_ycocg_r_to_rgb_bytes.argtypes = [p16, p16, p16, sz, p8, sz]
        L.alo_wavelet1d_init.argtypes = [C.POINTER(_W1D), C.c_int]
        for n in ("alo_wavelet1d_forward", "alo_wavelet1d_inverse"):
            getattr(L, n).argtypes = [C.POINTER(_W1D), p32, sz]
            getattr(L, n).restype = None
        for n in ("alo_wavelet2d_forward", "alo_wavelet2d_inverse"):
            getattr(L, n).argtypes = [C.POINTER(_W1D), p32, sz, sz]
            getattr(L, n).restype = None
        for n in ("alo_wavelet3d_forward", "alo_wavelet3d_inverse"):
            getattr(L, n).argtypes = [C.POINTER(_W1D), p32, sz, sz, sz]
            getattr(L, n).restype = None
        L.alo_quantize.argtypes = [C.c_int32, C.c_int32, C.c_int32]
        L.alo_quantize.restype = C.c_int32
        L.alo_dequantize.argtypes = [C.c_int32, C.c_int32]
        L.alo_dequantize.restype = C.c_int32
        L.alo_quantize_buffer.argtypes = [C.c_int32, C.c_int32, p32, sz, p32, sz]
        L.alo_dequantize_buffer.argtypes = [C.c_int32, p32, sz, p32, sz]
        L.alo_fastq_new.argtypes = [C.c_int32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
        L.alo_fastq_quantize_buffer.argtypes = [C.c_int32, C.c_int32, p32, sz, p32, sz]
        L.alo_rdo_bpp_from_quality.argtypes = [C.c_uint8]
        L.alo_rdo_bpp_from_quality.restype = C.c_double
        L.alo_rdo_estimate_variance.argtypes = [p32, sz]
        L.alo_rdo_estimate_variance.restype = C.c_double
        L.alo_rdo_compute_quantizer.argtypes = [C.c_double, p32, sz, C.c_int, p32, p32]
        L.alo_to_symbols.argtypes = [p32, sz, p8, sz]
        L.alo_from_symbols.argtypes = [p8, sz, p32, sz]
        L.alo_build_histogram.argtypes = [p8, sz, C.POINTER(C.c_uint32)]
        L.alo_build_histogram.restype = None
        L.alo_freq_table_uniform.argtypes = [C.c_uint32, C.POINTER(FreqTable)]
        L.alo_freq_table_from_histogram.argtypes = [C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(FreqTable)]
        L.alo_rans_encode.argtypes = [p8, sz, C.POINTER(FreqTable), C.POINTER(p8), C.POINTER(sz)]
        L.alo_rans_decode.argtypes = [p8, sz, sz, C.POINTER(FreqTable), p8]
        L.alo_rans_encode_interleaved.argtypes = [p8, sz, C.POINTER(FreqTable), C.POINTER(p8), C.POINTER(sz)]
        L.alo_rans_decode_interleaved.argtypes = [p8, sz, sz, C.POINTER(FreqTable), p8]
        L.alo_free.argtypes = [C.c_void_p]
        L.alo_free.restype = None
        L.alo_quality_to_step.argtypes = [C.c_uint8]
        L.alo_quality_to_step.restype = C.c_int32
        L.alo_encode.argtypes = [C.c_uint8, C.c_int, p8, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                 C.POINTER(p8), C.POINTER(C.c_uint64), C.POINTER(p32), C.POINTER(p8)]
        L.alo_decode.argtypes = [p8, C.c_uint64, C.POINTER(p8), C.POINTER(C.c_uint64), C.POINTER(p8)]
        L.alo_psnr.argtypes = [p8, p8, sz]
        L.alo_psnr.restype = C.c_double
        L.alo_generate.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, p8]
        L.alo_generate.restype = None
        _lib = L
    return _lib


def _p8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _p16(a):
    return a.ctypes.data_as(C.POINTER(C.c_int16))


def _p32(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _chk(rc):
    if rc != OK:
        raise OracleError(rc)


def _w1d(kind: int) -> _W1D:
    w = _W1D()
    _chk(lib().alo_wavelet1d_init(C.byref(w), kind))
    return w


# ---------------------------------------------------------------- colour (color.rs:199,245)
def rgb_bytes_to_ycocg_r(rgb: np.ndarray):
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8).ravel()
    n = rgb.size // 3
    y, co, cg = (np.empty(n, np.int16) for _ in range(3))
    _chk(lib().alo_rgb_bytes_to_ycocg_r(_p8(rgb), rgb.size, _p16(y), _p16(co), _p16(cg), n))
    return y, co, cg


def ycocg_r_to_rgb_bytes(y, co, cg):
    y, co, cg = (np.ascontiguousarray(a, dtype=np.int16).ravel() for a in (y, co, cg))
    out = np.empty(y.size * 3, np.uint8)
    _chk(lib().alo_ycocg_r_to_rgb_bytes(_p16(y), _p16(co), _p16(cg), y.size, _p8(out), out.size))
    return out


# ---------------------------------------------------------------- wavelet (wavelet.rs)
def wavelet1d_forward(kind, data):
    a = np.array(data, dtype=np.int32).ravel()
    lib().alo_wavelet1d_forward(C.byref(_w1d(kind)), _p32(a), a.size)
    return a


def wavelet1d_inverse(kind, data):
    a = np.array(data, dtype=np.int32).ravel()
    lib().alo_wavelet1d_inverse(C.byref(_w1d(kind)), _p32(a), a.size)
    return a


def wavelet2d_forward(kind, img, width, height):
    a = np.array(img, dtype=np.int32).ravel()
    assert a.size == width * height
    lib().alo_wavelet2d_forward(C.byref(_w1d(kind)), _p32(a), width, height)
    return a


def wavelet2d_inverse(kind, img, width, height):
    a = np.array(img, dtype=np.int32).ravel()
    assert a.size == width * height
    lib().alo_wavelet2d_inverse(C.byref(_w1d(kind)), _p32(a), width, height)
    return a


def wavelet3d_forward(kind, vol, width, height, depth):
    a = np.array(vol, dtype=np.int32).ravel()
    assert a.size == width * height * depth
    lib().alo_wavelet3d_forward(C.byref(_w1d(kind)), _p32(a), width, height, depth)
    return a


def wavelet3d_inverse(kind, vol, width, height, depth):
    a = np.array(vol, dtype=np.int32).ravel()
    assert a.size == width * height * depth
    lib().alo_wavelet3d_inverse(C.byref(_w1d(kind)), _p32(a), width, height, depth)
    return a


# ---------------------------------------------------------------- quant (quant.rs)
def quantize(step, dz, v):
    return lib().alo_quantize(step, dz, v)


def dequantize(step, q):
    return lib().alo_dequantize(step, q)


def quantize_buffer(step, dz, data):
    a = np.ascontiguousarray(data, dtype=np.int32).ravel()
    out = np.empty_like(a)
    _chk(lib().alo_quantize_buffer(step, dz, _p32(a), a.size, _p32(out), out.size))
    return out


def dequantize_buffer(step, data):
    a = np.ascontiguousarray(data, dtype=np.int32).ravel()
    out = np.empty_like(a)
    _chk(lib().alo_dequantize_buffer(step, _p32(a), a.size, _p32(out), out.size))
    return out


def fast_quantize_buffer(step, dz, data):
    a = np.ascontiguousarray(data, dtype=np.int32).ravel()
    out = np.empty_like(a)
    _chk(lib().alo_fastq_quantize_buffer(step, dz, _p32(a), a.size, _p32(out), out.size))
    return out


def fastq_new(step):
    r, s = C.c_uint64(), C.c_uint32()
    _chk(lib().alo_fastq_new(step, C.byref(r), C.byref(s)))
    return r.value, s.value


def rdo_bpp_from_quality(q):
    return lib().alo_rdo_bpp_from_quality(q)


def rdo_estimate_variance(coeffs):
    a = np.ascontiguousarray(coeffs, dtype=np.int32).ravel()
    return lib().alo_rdo_estimate_variance(_p32(a), a.size)


def rdo_compute_quantizer(target_bpp, coeffs, subband):
    a = np.ascontiguousarray(coeffs, dtype=np.int32).ravel()
    s, d = C.c_int32(), C.c_int32()
    _chk(lib().alo_rdo_compute_quantizer(target_bpp, _p32(a), a.size, subband, C.byref(s), C.byref(d)))
    return s.value, d.value


def to_symbols(coeffs):
    a = np.ascontiguousarray(coeffs, dtype=np.int32).ravel()
    out = np.empty(a.size, np.uint8)
    _chk(lib().alo_to_symbols(_p32(a), a.size, _p8(out), out.size))
    return out


def from_symbols(symbols):
    a = np.ascontiguousarray(symbols, dtype=np.uint8).ravel()
    out = np.empty(a.size, np.int32)
    _chk(lib().alo_from_symbols(_p8(a), a.size, _p32(out), out.size))
    return out


def build_histogram(symbols):
    a = np.ascontiguousarray(symbols, dtype=np.uint8).ravel()
    h = np.zeros(256, np.uint32)
    lib().alo_build_histogram(_p8(a), a.size, h.ctypes.data_as(C.POINTER(C.c_uint32)))
    return h


# ---------------------------------------------------------------- rANS (rans.rs)
def freq_table_from_histogram(hist) -> FreqTable:
    h = np.ascontiguousarray(hist, dtype=np.uint32).ravel()
    t = FreqTable()
    _chk(lib().alo_freq_table_from_histogram(h.ctypes.data_as(C.POINTER(C.c_uint32)), h.size, C.byref(t)))
    return t


def freq_table_uniform(n) -> FreqTable:
    t = FreqTable()
    _chk(lib().alo_freq_table_uniform(n, C.byref(t)))
    return t


def rans_encode(symbols, table: FreqTable) -> bytes:
    a = np.ascontiguousarray(symbols, dtype=np.uint8).ravel()
    out = C.POINTER(C.c_uint8)()
    n = C.c_size_t()
    _chk(lib().alo_rans_encode(_p8(a), a.size, C.byref(table), C.byref(out), C.byref(n)))
    try:
        return C.string_at(out, n.value)
    finally:
        lib().alo_free(out)


def rans_decode(stream: bytes, n: int, table: FreqTable) -> np.ndarray:
    buf = np.frombuffer(bytes(stream), dtype=np.uint8) if len(stream) else np.zeros(0, np.uint8)
    buf = np.ascontiguousarray(buf)
    out = np.empty(n, np.uint8)
    _chk(lib().alo_rans_decode(_p8(buf) if buf.size else None, buf.size, n, C.byref(table), _p8(out)))
    return out


def rans_encode_interleaved(symbols, table: FreqTable) -> bytes:
    """InterleavedRansEncoder::encode + finish (rans.rs:393-459)."""
    a = np.ascontiguousarray(symbols, dtype=np.uint8).ravel()
    out = C.POINTER(C.c_uint8)()
    n = C.c_size_t()
    _chk(lib().alo_rans_encode_interleaved(_p8(a) if a.size else None, a.size, C.byref(table), C.byref(out), C.byref(n)))
    try:
        return C.string_at(out, n.value)
    finally:
        lib().alo_free(out)


def rans_decode_interleaved(stream: bytes, n: int, table: FreqTable) -> np.ndarray:
    """InterleavedRansDecoder::new + decode_n (rans.rs:465-524)."""
    buf = np.ascontiguousarray(np.frombuffer(bytes(stream), dtype=np.uint8)) if len(stream) else np.zeros(0, np.uint8)
    out = np.empty(n, np.uint8)
    _chk(lib().alo_rans_decode_interleaved(_p8(buf) if buf.size else None, buf.size, n, C.byref(table), _p8(out)))
    return out


# ---------------------------------------------------------------- pipeline (pipeline.rs)
def quality_to_step(q):
    return lib().alo_quality_to_step(q)


def padded_dims(w, h, f):
    return w + (w & 1), h + (h & 1), (2 if f == 1 else f + (f & 1))


def encode(rgb, width, height, frames, quality=90, wavelet=CDF53, stages=False):
    """FrameEncoder::with_wavelet(quality, wavelet).encode(..).to_bytes()
    (pipeline.rs:377-507, :200-226).  stages=True also returns per-channel i32 coefficient
    volumes and u8 symbol planes (padded dims)."""
    a = np.ascontiguousarray(rgb, dtype=np.uint8).ravel()
    out = C.POINTER(C.c_uint8)()
    n = C.c_uint64()
    co = (C.POINTER(C.c_int32) * 3)()
    so = (C.POINTER(C.c_uint8) * 3)()
    rc = lib().alo_encode(quality, wavelet, _p8(a) if a.size else None, a.size, width, height, frames,
                          C.byref(out), C.byref(n), co if stages else None, so if stages else None)
    if rc != OK:
        raise OracleError(rc)
    try:
        alc = C.string_at(out, n.value)
    finally:
        lib().alo_free(out)
    if not stages:
        return alc
    pw, ph, pf = padded_dims(width, height, frames)
    N = pw * ph * pf if width * height * frames else 0
    coeffs, syms = [], []
    for c in range(3):
        if N and co[c]:
            coeffs.append(np.ctypeslib.as_array(co[c], shape=(N,)).copy())
            syms.append(np.ctypeslib.as_array(so[c], shape=(N,)).copy())
            lib().alo_free(co[c])
            lib().alo_free(so[c])
        else:
            coeffs.append(np.zeros(0, np.int32))
            syms.append(np.zeros(0, np.uint8))
    return alc, coeffs, syms


def decode(alc: bytes, stages=False):
    """EncodedChunk::from_bytes + FrameDecoder::decode (pipeline.rs:235-313, :537-624)."""
    buf = np.frombuffer(bytes(alc), dtype=np.uint8)
    out = C.POINTER(C.c_uint8)()
    n = C.c_uint64()
    so = (C.POINTER(C.c_uint8) * 3)()
    rc = lib().alo_decode(_p8(buf) if buf.size else None, buf.size, C.byref(out), C.byref(n),
                          so if stages else None)
    if rc != OK:
        raise OracleError(rc)
    try:
        rgb = np.frombuffer(C.string_at(out, n.value), dtype=np.uint8).copy()
    finally:
        lib().alo_free(out)
    if not stages:
        return rgb
    syms = []
    nsym = int.from_bytes(alc[18 + 12:18 + 16], "little")
    for c in range(3):
        if so[c]:
            syms.append(np.ctypeslib.as_array(so[c], shape=(nsym,)).copy())
            lib().alo_free(so[c])
        else:
            syms.append(np.zeros(0, np.uint8))
    return rgb, syms


def psnr(a, b):
    a = np.ascontiguousarray(a, dtype=np.uint8).ravel()
    b = np.ascontiguousarray(b, dtype=np.uint8).ravel()
    assert a.size == b.size
    return lib().alo_psnr(_p8(a), _p8(b), a.size)


# ---------------------------------------------------------------- synthetic inputs
G0, G1, G2 = 0, 1, 2
SEED = 0x5EED0001


def generate(kind, w, h, f, seed=SEED):
    out = np.empty(w * h * f * 3, np.uint8)
    lib().alo_generate(kind, seed, w, h, f, _p8(out))
    return out
