"""Front-end (and back-end) time per 1920x1080x64 chunk for the three wavelets: CUDA-event stage times of the batch API,
one chunk alone and a batch of `--batch` chunks in one launch, with the .alc / decoded digests checked against the
committed oracle digests where they exist (BASELINE configs 1 and 2).

    python tools/frontend_time.py [--batch 8] [--reps 3] [--decode]
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--decode", action="store_true")
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--lib", default=None, help="experiment build of libalice_codec")
a = ap.parse_args()
pkg = load_package()
api = pkg.Api(a.lib) if a.lib else pkg.default_api()
api.set_device(0)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
gold = json.load(open(os.path.join(ROOT, "tests", "golden", "fullsize.json")))
W, H, F = a.width, a.height, 64
n = W * H * F * 3
st = torch.cuda.current_stream()
bufs = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(a.batch)]
outs = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(a.batch if a.decode else 0)]
for i, t in enumerate(bufs):
    api._chk(api.lib.alice_codec_synth_rgb_device(1, 0x5EED0001 + i, W, H, F, C.c_void_p(t.data_ptr()), C.c_void_p(st.cuda_stream)))
cases = [("cdf53", 90, "cfg1_cdf53_q90_1080p64"), ("cdf97", 80, "cfg2_cdf97_q80_1080p64"), ("haar", 75, None)]
if (W, H) == (3840, 2160):
    cases = [("haar", 75, "cfg3_haar_q75_4k64"), ("cdf97", 80, "cfg5_cdf97_q80_4k64_chunk0")]
for wavelet, q, gname in cases:
    for B in sorted({1, a.batch}):
        b = pkg.ChunkBatch(q, wavelet, W, H, F, B, stream=st.cuda_stream, api=api)
        fe, be = [], []
        for _ in range(a.reps):
            b.encode_device([t.data_ptr() for t in bufs[:B]])
            fe.append(b.timings()[0] / B)
            if a.decode:
                b.decode_device([t.data_ptr() for t in outs[:B]])
                be.append(b.timings()[5] / B)
        rec = {"wavelet": wavelet, "quality": q, "shape": [W, H, F], "chunks_per_launch": B,
               "frontend_ms_per_chunk": [round(x, 4) for x in fe], "best_ms": round(min(fe), 4),
               "alg_gb_s": round(6 * W * H * F / min(fe) / 1e6, 1), "frac_of_6464.9": round(6 * W * H * F / min(fe) / 1e6 / 6464.9, 4)}
        if a.decode:
            rec["backend_ms_per_chunk"] = [round(x, 4) for x in be]
        if gname:
            rec["alc_equals_oracle_digest"] = hashlib.sha256(b.get_chunk(0).to_bytes()).hexdigest() == gold[gname]["sha256_alc"]
            if a.decode:
                rec["decoded_equals_oracle_digest"] = hashlib.sha256(outs[0].cpu().numpy().tobytes()).hexdigest() == gold[gname]["sha256_decoded"]
        print(json.dumps(rec), flush=True)
        b.close()
