"""Generates tests/golden/fullsize.json: sha256 digests of the ORACLE's outputs at BASELINE.json's full sizes.

The reference is a Rust crate and cannot run in this image (no rustc/cargo), so these are digests of the CPU
oracle (oracle/alice_oracle.c), which is itself pinned by the reference's exact-value tests and the SURVEY
Appendix C vectors (tests/test_oracle.py).  Run on a CPU box:  python tests/golden/make_golden.py [names...]
The -m gpu tests compare the CUDA path's digests with these without re-running the oracle at full size.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fullsize.json")
CASES = {
    # name: (kind, w, h, f, quality, wavelet, seed)        BASELINE.json configs
    "cfg1_cdf53_q90_1080p64": (O.G1, 1920, 1080, 64, 90, 0, O.SEED),
    "cfg2_cdf97_q80_1080p64": (O.G1, 1920, 1080, 64, 80, 1, O.SEED),
    "cfg3_haar_q75_4k64": (O.G1, 3840, 2160, 64, 75, 2, O.SEED),
    "cfg5_cdf97_q80_4k64_chunk0": (O.G1, 3840, 2160, 64, 80, 1, O.SEED),
    "cfg5_cdf97_q80_4k64_chunk1": (O.G1, 3840, 2160, 64, 80, 1, O.SEED + 1),
    "noise_cdf53_q90_1080p8": (O.G2, 1920, 1080, 8, 90, 0, O.SEED),
    "odd_cdf97_q80_1919x1079x63": (O.G1, 1919, 1079, 63, 80, 1, O.SEED),
}


def sha(b):
    return hashlib.sha256(b).hexdigest()


def main():
    names = sys.argv[1:] or list(CASES)
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name in names:
        kind, w, h, f, q, wv, seed = CASES[name]
        t0 = time.time()
        rgb = O.generate(kind, w, h, f, seed)
        t1 = time.time()
        alc, coeffs, syms = O.encode(rgb, w, h, f, q, wv, stages=True)
        t2 = time.time()
        dec = O.decode(alc)
        t3 = time.time()
        lens = [int.from_bytes(alc[18 + c * 1040:18 + c * 1040 + 4], "little") for c in range(3)]
        res[name] = {
            "kind": kind, "w": w, "h": h, "f": f, "quality": q, "wavelet": wv, "seed": seed,
            "sha256_rgb_in": sha(rgb.tobytes()), "alc_len": len(alc), "stream_lens": lens,
            "sha256_alc": sha(alc), "sha256_header": sha(alc[:3138]),
            "sha256_coeffs": [sha(c.tobytes()) for c in coeffs], "sha256_symbols": [sha(s.tobytes()) for s in syms],
            "sha256_decoded": sha(dec.tobytes()),
            "oracle_seconds": {"generate": round(t1 - t0, 2), "encode": round(t2 - t1, 2), "decode": round(t3 - t2, 2)},
        }
        print(name, res[name]["alc_len"], res[name]["oracle_seconds"], flush=True)
        del rgb, alc, coeffs, syms, dec
        json.dump(res, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
