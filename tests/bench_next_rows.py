"""Measurements for the SURVEY.md 8(f) rows beside the hot path (RDO statistics, device PSNR, interleaved rANS),
each through the C ABI with HOST buffers (copies included), next to the C oracle on one host core.
Prints one JSON line per row; run on a GPU box:  python tests/bench_next_rows.py > gpurun_out/next_rows.jsonl"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tests/ -> repo root
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
from oracle import oracle as O  # noqa: E402  (the checker and the CPU baseline leg)

pkg = load_package()
api = pkg.default_api()
api.set_device(0)


def best(f, reps=3):
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        r = f()
        ts.append(time.perf_counter() - t)
    return min(ts), r


def octants(vol, w, h, d):
    v = vol.reshape(d, h, w)
    hx, hy, ht = w // 2, h // 2, d // 2
    return [np.ascontiguousarray(v[(ht if sb & 1 else 0):(ht if sb & 1 else 0) + ht,
                                   (hy if sb & 2 else 0):(hy if sb & 2 else 0) + hy,
                                   (hx if sb & 4 else 0):(hx if sb & 4 else 0) + hx]).reshape(-1) for sb in range(8)]


def row_rdo():
    w, h, d = 960, 540, 64
    y = O.rgb_bytes_to_ycocg_r(O.generate(O.G1, w, h, d))[0].astype(np.int32)
    vol = api.wavelet3d(1, y, w, h, d)
    bpp = O.rdo_bpp_from_quality(80)
    api.rdo_compute_all_quantizers(bpp, vol[:8 * 8 * 8], 8, 8, 8)          # warm-up
    tg, got = best(lambda: api.rdo_compute_all_quantizers(bpp, vol, w, h, d))
    tq, (q, got2) = best(lambda: api.rdo_quantize_volume(bpp, vol, w, h, d))
    import torch
    dvol = torch.from_numpy(vol).cuda()
    torch.cuda.synchronize()
    td, got3 = best(lambda: api.rdo_compute_all_quantizers_device(bpp, dvol.data_ptr(), w, h, d), reps=5)
    t0 = time.perf_counter()
    octs = octants(vol, w, h, d)
    want = [O.rdo_compute_quantizer(bpp, o, sb) for sb, o in enumerate(octs)]
    tc = time.perf_counter() - t0
    n = w * h * d
    print(json.dumps({"row": "8f-2 AnalyticalRDO::compute_all_quantizers + per-octant FastQuantizer",
                      "volume": f"{w}x{h}x{d} CDF 9/7 coefficients of G1 luma ({n} i32)",
                      "bit_exact_vs_oracle": got == want and got2 == want and got3 == want,
                      "gpu_stats_device_resident_s": round(td, 5), "gpu_Mcoef_s_stats_device_resident": round(n / td / 1e6, 1),
                      "gpu_stats_s": round(tg, 4), "gpu_stats_plus_quantise_s": round(tq, 4),
                      "gpu_Mcoef_s_stats": round(n / tg / 1e6, 1), "gpu_Mcoef_s_fused": round(n / tq / 1e6, 1),
                      "cpu_oracle_stats_s": round(tc, 3), "cpu_Mcoef_s_stats": round(n / tc / 1e6, 1), "cpu_cores": 1,
                      "note": "host buffers: H2D of the volume (and D2H of the quantised volume) inside the timed call"}),
          flush=True)


def row_psnr():
    import torch
    n = 3 * 1920 * 1080 * 64
    a = torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda")
    b = (a.to(torch.int16) + torch.randint(-3, 4, (n,), dtype=torch.int16, device="cuda")).clamp(0, 255).to(torch.uint8)
    api.psnr_device(a.data_ptr(), b.data_ptr(), n)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    st = torch.cuda.current_stream().cuda_stream
    ev[0].record()
    reps = 10
    for _ in range(reps):
        v = api.psnr_device(a.data_ptr(), b.data_ptr(), n, st)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / reps
    sample = 3 * 1920 * 1080 * 2
    ha, hb = a[:sample].cpu().numpy(), b[:sample].cpu().numpy()
    tc, vc = best(lambda: O.psnr(ha, hb), reps=2)
    print(json.dumps({"row": "8f-3 alice_codec_psnr on device buffers", "bytes_per_call": 2 * n, "ms_per_call": round(ms, 4),
                      "gb_s": round(2 * n / ms / 1e6, 1), "psnr_db": v,
                      "equals_host_function_on_sample": api.psnr_device(a.data_ptr(), b.data_ptr(), sample) == vc,
                      "cpu_oracle_gb_s": round(2 * sample / tc / 1e9, 2), "cpu_cores": 1}), flush=True)


def row_interleaved():
    n = 1 << 25
    rng = np.random.default_rng(1)
    sy = np.minimum(rng.geometric(0.3, n) - 1, 255).astype(np.uint8)
    hist = np.bincount(sy, minlength=256).astype(np.uint32)
    api.rans_encode_interleaved(sy[:4096], hist)
    te, blob = best(lambda: api.rans_encode_interleaved(sy, hist), reps=2)
    td, back = best(lambda: api.rans_decode_interleaved(blob, n, hist), reps=2)
    te1, blob1 = best(lambda: api.rans_encode(sy, hist), reps=1)
    td1, back1 = best(lambda: api.rans_decode(blob1, n, hist), reps=1)
    m = 1 << 22
    t = O.freq_table_from_histogram(hist)
    tce, cblob = best(lambda: O.rans_encode_interleaved(sy[:m], t), reps=1)
    tcd, _ = best(lambda: O.rans_decode_interleaved(cblob, m, t), reps=1)
    # (decoding need not return the input: from_histogram tables are routinely malformed, SURVEY.md 0.7 — the check is
    # equality with the oracle's decode of the same container)
    dec_ok = bool(np.array_equal(api.rans_decode_interleaved(cblob, m, hist), O.rans_decode_interleaved(cblob, m, t)))
    print(json.dumps({"row": "8f-4 InterleavedRansEncoder/Decoder (4 streams, not .alc)", "symbols": n,
                      "decode_equals_oracle_on_sample": dec_ok,
                      "container_equals_oracle_on_sample": api.rans_encode_interleaved(sy[:m], hist) == cblob,
                      "gpu_encode_Msym_s": round(n / te / 1e6, 1), "gpu_decode_Msym_s": round(n / td / 1e6, 1),
                      "gpu_single_stream_encode_Msym_s": round(n / te1 / 1e6, 1),
                      "gpu_single_stream_decode_Msym_s": round(n / td1 / 1e6, 1),
                      "cpu_oracle_encode_Msym_s": round(m / tce / 1e6, 1), "cpu_oracle_decode_Msym_s": round(m / tcd / 1e6, 1),
                      "cpu_cores": 1, "note": "host buffers; one call = 4 lanes, throughput scales with calls in flight"}),
          flush=True)


if __name__ == "__main__":
    for f in (row_rdo, row_psnr, row_interleaved):
        try:
            f()
        except Exception as e:  # keep the other rows
            print(json.dumps({"row": f.__name__, "error": repr(e)}), flush=True)
