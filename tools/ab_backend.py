"""A/B of experiment builds on the decode back-end (k_inv_t + k_inv_yx): one process, one 1920x1080x64 CDF 9/7 G1
chunk per build, CUDA-event time of the back-end stage, decoded RGB compared on the device with the first build's.

    python tools/ab_backend.py base inv_rolled          # names = lib/libalice_codec_<name>.so
"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
W, H, F = 1920, 1080, 64
REPS = int(os.environ.get("AB_REPS", "2"))
names = sys.argv[1:] or ["base"]
libdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "alice-codec_b200", "lib")
st = torch.cuda.current_stream()
n = W * H * F * 3
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
first = None
for name in names:
    path = os.path.join(libdir, "libalice_codec.so" if name == "base" else f"libalice_codec_{name}.so")
    try:
        api = pkg.Api(path)
        api.set_device(0)
        api._chk(api.lib.alice_codec_synth_rgb_device(1, 0x5EED0001, W, H, F, C.c_void_p(d_in.data_ptr()),
                                                      C.c_void_p(st.cuda_stream)))
        warm = pkg.ChunkBatch(80, "cdf97", 64, 36, F, 1, stream=st.cuda_stream, api=api)   # loads the same kernels
        warm.encode_device([d_in.data_ptr()])
        warm.decode_device([d_out.data_ptr()])
        warm.close()
        b = pkg.ChunkBatch(80, "cdf97", W, H, F, 1, stream=st.cuda_stream, api=api)
        b.encode_device([d_in.data_ptr()])
        be = []
        for _ in range(REPS):
            d_out.zero_()
            b.decode_device([d_out.data_ptr()])
            be.append(b.timings()[5])
        b.close()
        torch.cuda.synchronize()
        if first is None:
            first = d_out.clone()
        rec = {"build": name, "backend_ms": [round(x, 4) for x in be], "best_ms": round(min(be), 4),
               "alg_gb_s": round(6 * W * H * F / min(be) / 1e6, 1), "same_rgb_as_first": bool(torch.equal(first, d_out))}
        print(json.dumps(rec), flush=True)
        if os.environ.get("AB_GOLDEN"):   # the chunk is BASELINE config 2: compare with the committed oracle digest
            import hashlib
            gold = json.load(open(os.path.join(os.path.dirname(libdir), "..", "tests", "golden", "fullsize.json")))
            digest = hashlib.sha256(d_out.cpu().numpy().tobytes()).hexdigest()
            print(json.dumps({"build": name, "decoded_equals_oracle_digest": digest == gold["cfg2_cdf97_q80_1080p64"]["sha256_decoded"]}), flush=True)
    except Exception as e:
        print(json.dumps({"build": name, "error": repr(e)}), flush=True)
