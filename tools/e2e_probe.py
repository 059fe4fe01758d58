"""Probe of the host-buffer path: raw PCIe bandwidth with pinned memory, then wall-clock vs kernel time of
batch_encode_host / batch_decode_host for one batch."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
ap = argparse.ArgumentParser(); ap.add_argument("--chunks", type=int, default=8); a = ap.parse_args()
pkg = load_package(); api = pkg.default_api(); api.set_device(0)
W, H, F = 1920, 1080, 64
n = W * H * F * 3
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{name} pinned: {5 * n / dt / 1e9:.1f} GB/s")
st = torch.cuda.current_stream()
import ctypes as C
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
h_in = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(a.chunks)]
h_out = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(a.chunks)]
for i in range(a.chunks):
    api._chk(api.lib.alice_codec_synth_rgb_device(1, 0x5EED0001 + i, W, H, F, C.c_void_p(d_in.data_ptr()), C.c_void_p(st.cuda_stream)))
    h_in[i].copy_(d_in)
torch.cuda.synchronize()
b = pkg.ChunkBatch(80, "cdf97", W, H, F, a.chunks, stream=st.cuda_stream, api=api)
for it in range(2):
    t0 = time.perf_counter(); ck = b.encode_host([t.data_ptr() for t in h_in]); t1 = time.perf_counter()
    b.decode_host(ck, [t.data_ptr() for t in h_out]); t2 = time.perf_counter()
    ms = b.timings()
    print(f"it {it}: encode_host {t1 - t0:.3f} s (kernels {sum(ms[:3]) / 1e3:.3f}), decode_host {t2 - t1:.3f} s (kernels {sum(ms[3:6]) / 1e3:.3f}); "
          f"rgb bytes {a.chunks * n / 1e9:.2f} GB each way, payload {sum(c.compressed_size for c in ck) / 1e9:.2f} GB")
# component timing through the device-pointer API
d_ins = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(a.chunks)]
for it in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(a.chunks): d_ins[i].copy_(h_in[i], non_blocking=True)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    b.encode_device([t.data_ptr() for t in d_ins]); t2 = time.perf_counter()
    cks = [b.get_chunk(i) for i in range(a.chunks)]; t3 = time.perf_counter()
    del cks; t4 = time.perf_counter()
    print(f"components it {it}: H2D {t1 - t0:.3f} s, encode_device {t2 - t1:.3f} s (kernels {sum(b.timings()[:3]) / 1e3:.3f}), get_chunk x{a.chunks} {t3 - t2:.3f} s, destroy {t4 - t3:.3f} s")
