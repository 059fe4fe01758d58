"""rANS throughput against streams per SM: one process, 1920x1080x`--frames` G1 chunks (short streams so that many fit),
encode + decode of `chunks` chunks in place through the device-pointer batch API, for several chunk counts.
Prints one JSON line per (chunks, decoder layout): launch times and Msym/s per SM.

    python tools/rans_occupancy.py --frames 8 --chunks 99,197,296,395,592
"""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=8)
ap.add_argument("--chunks", default="99,197,296,395,592")
ap.add_argument("--quality", type=int, default=80)
ap.add_argument("--wavelet", default="cdf97")
ap.add_argument("--envs", default="", help="comma-separated NAME=VALUE settings tried one after another (besides none)")
ap.add_argument("--lib", default=None)
a = ap.parse_args()
pkg = load_package()
api = pkg.Api(a.lib) if a.lib else pkg.default_api()
api.set_device(0)
st = torch.cuda.current_stream()
W, H, F = 1920, 1080, a.frames
n = W * H * F * 3
n_sm = torch.cuda.get_device_properties(0).multi_processor_count
counts = [int(x) for x in a.chunks.split(",")]
bufs = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(max(counts))]


def synth(k):
    for i in range(k):
        api._chk(api.lib.alice_codec_synth_rgb_device(1, 0x5EED0001 + i, W, H, F, C.c_void_p(bufs[i].data_ptr()),
                                                      C.c_void_p(st.cuda_stream)))


envs = [None] + [e for e in a.envs.split(",") if e]
for env in envs:
    if env:
        k, v = env.split("=")
        os.environ[k] = v
    for B in counts:
        ptrs = [bufs[i].data_ptr() for i in range(B)]
        b = pkg.ChunkBatch(a.quality, a.wavelet, W, H, F, B, stream=st.cuda_stream, api=api, shared_workspace=True)
        best = None
        for rep in range(2):
            synth(B)
            b.encode_device(ptrs, ptrs)
            b.decode_device(ptrs)
            ms = b.timings()
            if best is None or ms[2] + ms[4] < best[2] + best[4]:
                best = ms
        b.close()
        nsym = W * H * F * 3 * B
        print(json.dumps({"env": env, "chunks": B, "streams": 3 * B, "streams_per_sm": round(3 * B / n_sm, 2),
                          "frames": F, "enc_ms": round(best[2], 2), "dec_ms": round(best[4], 2),
                          "enc_msym_s_per_sm": round(nsym / best[2] / 1e3 / n_sm, 1),
                          "dec_msym_s_per_sm": round(nsym / best[4] / 1e3 / n_sm, 1),
                          "enc_msym_s_per_lane_if_all_resident": round(W * H * F / best[2] / 1e3, 1),
                          "dec_msym_s_per_lane_if_all_resident": round(W * H * F / best[4] / 1e3, 1),
                          "frontend_ms_per_chunk": round(best[0] / B, 4), "backend_ms_per_chunk": round(best[5] / B, 4)}),
              flush=True)
    if env:
        del os.environ[env.split("=")[0]]
