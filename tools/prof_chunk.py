"""Small driver for ncu captures: one warm-up and one measured encode+decode of `--chunks` G1 chunks of
1920x1080x`--frames` through the device-pointer batch API (same kernels as bench.py, shorter rANS streams)."""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=4)
ap.add_argument("--chunks", type=int, default=2)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--quality", type=int, default=80)
ap.add_argument("--wavelet", default="cdf97")
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--lib", default=None, help="experiment build of the library")
ap.add_argument("--skip-rans", action="store_true", help="front-end / back-end kernels only (stage API)")
a = ap.parse_args()
pkg = load_package()
api = pkg.Api(a.lib) if a.lib else pkg.default_api()
api.set_device(0)
st = torch.cuda.current_stream()
n = a.width * a.height * a.frames * 3
d_in = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(a.chunks)]
d_out = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(a.chunks)]
for i, t in enumerate(d_in):
    api._chk(api.lib.alice_codec_synth_rgb_device(1, 0x5EED0001 + i, a.width, a.height, a.frames,
                                                  C.c_void_p(t.data_ptr()), C.c_void_p(st.cuda_stream)))
b = pkg.ChunkBatch(a.quality, a.wavelet, a.width, a.height, a.frames, a.chunks, stream=st.cuda_stream, api=api)
for _ in range(1 + a.reps):
    b.encode_device([t.data_ptr() for t in d_in])
    b.decode_device([t.data_ptr() for t in d_out])
torch.cuda.synchronize()
ms = b.timings()
npx = a.width * a.height * a.frames
print("stage ms: frontend %.3f tables %.3f rans_enc %.3f | tables %.3f rans_dec %.3f backend %.3f" % tuple(ms[:6]))
print("rans enc %.1f Msym/s/lane, dec %.1f Msym/s/lane; frontend %.1f GB/s alg, backend %.1f GB/s alg" % (
    npx / ms[2] / 1e3, npx / ms[4] / 1e3, 6 * npx * a.chunks / ms[0] / 1e6, 6 * npx * a.chunks / ms[5] / 1e6))
