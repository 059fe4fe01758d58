"""Standalone Wavelet3D / Wavelet2D forward + inverse on a device-resident i32 volume (alice_codec_wavelet3d_device /
_wavelet2d_device), CUDA-event time against the 8 B/sample algorithmic roofline of SURVEY.md 8(d), results checked
against the oracle on a small volume of the same code path.

    python tools/wavelet_time.py [--width 1920 --height 1080 --depth 64]
"""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402
import oracle as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--depth", type=int, default=64)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
pkg = load_package()
api = pkg.default_api()
api.set_device(0)
L = api.lib
st = torch.cuda.current_stream()
peak = 6464.9
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = float(json.load(open(pk))["hbm_gbs"])

# parity of the device entry points on a small volume of the same (fast) code path
w, h, d = 128, 36, 10
x = np.random.default_rng(5).integers(-(1 << 20), 1 << 20, w * h * d, dtype=np.int64).astype(np.int32)
for wv in (0, 1, 2):
    dx, dt = torch.from_numpy(x.copy()).cuda(), torch.empty(w * h * d, dtype=torch.int32, device="cuda")
    api._chk(L.alice_codec_wavelet3d_device(wv, 0, C.c_void_p(dx.data_ptr()), C.c_void_p(dt.data_ptr()), w, h, d, C.c_void_p(st.cuda_stream)))
    assert np.array_equal(dx.cpu().numpy(), O.wavelet3d_forward(wv, x, w, h, d)), wv
    api._chk(L.alice_codec_wavelet3d_device(wv, 1, C.c_void_p(dx.data_ptr()), C.c_void_p(dt.data_ptr()), w, h, d, C.c_void_p(st.cuda_stream)))
    assert np.array_equal(dx.cpu().numpy(), O.wavelet3d_inverse(wv, O.wavelet3d_forward(wv, x, w, h, d), w, h, d)), wv

W, H, D = a.width, a.height, a.depth
n = W * H * D
vol = torch.randint(-3000, 3000, (n,), dtype=torch.int32, device="cuda")
tmp = torch.empty_like(vol)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for wv, name in ((0, "cdf53"), (1, "cdf97"), (2, "haar")):
    rec = {"wavelet": name, "shape": [W, H, D], "algorithmic_bytes": 8 * n}
    for inv, key in ((0, "forward3d"), (1, "inverse3d")):
        best = 1e9
        for _ in range(a.reps):
            ev[0].record(st)
            api._chk(L.alice_codec_wavelet3d_device(wv, inv, C.c_void_p(vol.data_ptr()), C.c_void_p(tmp.data_ptr()), W, H, D, C.c_void_p(st.cuda_stream)))
            ev[1].record(st)
            torch.cuda.synchronize()
            best = min(best, ev[0].elapsed_time(ev[1]))
        rec[key + "_ms"] = round(best, 4)
        rec[key + "_frac_of_8B_roofline"] = round(8 * n / best / 1e6 / peak, 4)
    for inv, key in ((0, "forward2d_x64frames"), (1, "inverse2d_x64frames")):
        best = 1e9
        for _ in range(a.reps):
            ev[0].record(st)
            api._chk(L.alice_codec_wavelet2d_device(wv, inv, C.c_void_p(vol.data_ptr()), C.c_void_p(tmp.data_ptr()), W, H, D, C.c_void_p(st.cuda_stream)))
            ev[1].record(st)
            torch.cuda.synchronize()
            best = min(best, ev[0].elapsed_time(ev[1]))
        rec[key + "_ms"] = round(best, 4)
        rec[key + "_frac_of_8B_roofline"] = round(8 * n / best / 1e6 / peak, 4)
    print(json.dumps(rec), flush=True)
