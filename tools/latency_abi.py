"""What a Unity / UE5 / PyO3 consumer of the 20-symbol reference ABI gets for ONE chunk (src/ffi.rs:116-161):
alice_codec_encode + alice_codec_chunk_to_bytes + alice_codec_chunk_from_bytes + alice_codec_decode of one
1920x1080x64 chunk held in pageable host memory, wall clock per call, next to the C oracle on one host core.

    python tools/latency_abi.py [--reps 3] [--no-oracle]
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402
import oracle as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--no-oracle", action="store_true")
a = ap.parse_args()
pkg = load_package()
api = pkg.default_api()
api.set_device(0)
L = api.lib
W, H, F, Q = 1920, 1080, 64, 90           # alice_codec_encoder_create is always CDF 5/3 (ffi.rs:92): BASELINE config 1
gold = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "fullsize.json")))["cfg1_cdf53_q90_1080p64"]
rgb = O.generate(O.G1, W, H, F)            # pageable numpy memory
u8p = C.POINTER(C.c_uint8)
best = None
for rep in range(a.reps + 1):
    t0 = time.perf_counter()
    enc = L.alice_codec_encoder_create(Q)
    ck = L.alice_codec_encode(enc, rgb.ctypes.data_as(u8p), rgb.size, W, H, F)
    t1 = time.perf_counter()
    n = C.c_uint32()
    p = L.alice_codec_chunk_to_bytes(ck, C.byref(n))
    alc = C.string_at(p, n.value)
    L.alice_codec_data_free(p, n.value)
    L.alice_codec_chunk_destroy(ck)
    L.alice_codec_encoder_destroy(enc)
    t2 = time.perf_counter()
    buf = np.frombuffer(alc, dtype=np.uint8)
    ck2 = L.alice_codec_chunk_from_bytes(buf.ctypes.data_as(u8p), buf.size)
    t3 = time.perf_counter()
    m = C.c_uint32()
    q = L.alice_codec_decode(ck2, C.byref(m))
    t4 = time.perf_counter()
    out = np.empty(m.value, np.uint8)
    C.memmove(out.ctypes.data, q, m.value)
    L.alice_codec_data_free(q, m.value)
    L.alice_codec_chunk_destroy(ck2)
    rec = {"encode_s": t1 - t0, "to_bytes_s": t2 - t1, "from_bytes_s": t3 - t2, "decode_s": t4 - t3, "total_s": t4 - t0}
    if rep > 0 and (best is None or rec["total_s"] < best["total_s"]):   # rep 0 = warm-up (context, engine pool)
        best = rec
ok = hashlib.sha256(alc).hexdigest() == gold["sha256_alc"] and hashlib.sha256(out.tobytes()).hexdigest() == gold["sha256_decoded"]
line = {"what": "one 1920x1080x64 chunk, CDF 5/3 q=90, through the reference's 20-symbol C ABI from pageable host memory",
        "gpu": {k: round(v, 3) for k, v in best.items()}, "gpu_frames_per_s": round(F / best["total_s"], 2),
        "bit_exact_vs_oracle_digest": ok, "alc_bytes": len(alc)}
if not a.no_oracle:
    t0 = time.perf_counter()
    ralc = O.encode(rgb, W, H, F, Q, 0)
    t1 = time.perf_counter()
    O.decode(ralc)
    t2 = time.perf_counter()
    line["oracle_one_core"] = {"encode_s": round(t1 - t0, 2), "decode_s": round(t2 - t1, 2), "total_s": round(t2 - t0, 2)}
    line["oracle_frames_per_s"] = round(F / (t2 - t0), 2)
print(json.dumps(line))
