"""A/B of experiment builds of the library on the encode front-end (k_fwd_xy + k_fwd_t_quant): one process, one
1920x1080x64 CDF 9/7 G1 chunk per build, CUDA-event time of the front-end stage (rANS runs too but is not compared),
and the .alc digest of every build against the product build.

    python tools/ab_frontend.py base t_rolled t_rolled3 xyu2 xyu4      # names = lib/libalice_codec_<name>.so
"""
import ctypes as C
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
W, H, F, REPS = 1920, 1080, 64, 3
names = sys.argv[1:] or ["base"]
libdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "alice-codec_b200", "lib")
st = torch.cuda.current_stream()
d_in = torch.empty(W * H * F * 3, dtype=torch.uint8, device="cuda")
base_digest = None
for name in names:
    path = os.path.join(libdir, "libalice_codec.so" if name == "base" else f"libalice_codec_{name}.so")
    try:
        api = pkg.Api(path)
        api.set_device(0)
        api._chk(api.lib.alice_codec_synth_rgb_device(1, 0x5EED0001, W, H, F, C.c_void_p(d_in.data_ptr()),
                                                      C.c_void_p(st.cuda_stream)))
        warm = pkg.ChunkBatch(80, "cdf97", 64, 36, F, 1, stream=st.cuda_stream, api=api)   # loads the same kernels
        warm.encode_device([d_in.data_ptr()])
        warm.close()
        b = pkg.ChunkBatch(80, "cdf97", W, H, F, 1, stream=st.cuda_stream, api=api)
        fe = []
        for _ in range(REPS):
            b.encode_device([d_in.data_ptr()])
            fe.append(b.timings()[0])
        digest = hashlib.sha256(b.get_chunk(0).to_bytes()).hexdigest()
        b.close()
        if base_digest is None:
            base_digest = digest
        print(json.dumps({"build": name, "frontend_ms": [round(x, 4) for x in fe], "best_ms": round(min(fe), 4),
                          "alg_gb_s": round(6 * W * H * F / min(fe) / 1e6, 1), "alc_sha256": digest[:16],
                          "same_alc_as_first": digest == base_digest}), flush=True)
    except Exception as e:
        print(json.dumps({"build": name, "error": repr(e)}), flush=True)
