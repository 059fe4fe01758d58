"""Hardware parity spot-check of an experiment build before it becomes the default (not collected by pytest):
    python tests/check_variant_gpu.py alice-codec_b200/lib/libalice_codec_<name>.so
Runs the 64-frame encode/decode parity checks for every wavelet and the foreign-header decodes (narrow and wide
arithmetic) against the oracle; prints one JSON line."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(HERE), HERE]       # repo root (oracle/, __graft_entry__) and tests/ (parity)
import oracle as O  # noqa: E402
import parity  # noqa: E402

api = parity.pkg.Api(sys.argv[1])
api.set_device(0)
try:
    for wt in (0, 1, 2):
        for shape in ((260, 4, 64), (124, 62, 64), (16, 6, 63)):
            parity.check_encode_decode(api, O.G1, *shape, 80, wt)
        parity.check_encode_decode(api, O.G2, 64, 8, 64, 100, wt)
        parity.check_encode_decode(api, O.G0, 64, 8, 64, 0, wt)
    parity.check_decode_foreign_headers(api, np.random.default_rng(3))
    print(json.dumps({"lib": os.path.basename(sys.argv[1]), "parity": "ok"}), flush=True)
except Exception as e:
    print(json.dumps({"lib": os.path.basename(sys.argv[1]), "parity": "FAILED", "error": repr(e)[:400]}), flush=True)
