#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's config, measured on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--chunks B] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[1] — CDF 9/7, q=80, 1920x1080x64 RGB chunks (synthetic G1
volumes, SURVEY.md Appendix D, one seed per chunk), encode + decode round trip.  One *step* = one pass of the
hot path over a batch of B independent chunks per GPU:
    encode : RGB -> YCoCg-R -> 3-D lifting -> quantise -> symbols + histograms -> frequency tables -> 3B rANS lanes
    decode : tables -> 3B rANS lanes -> dequantise -> inverse lifting -> RGB
`value`  = frames/s with the RGB inputs already resident in HBM (whole job, all GPUs).
`e2e`    = the same metric through the C-ABI batch calls with HOST buffers (pinned): H2D of every RGB chunk,
           D2H of the .alc payloads, H2D of the payloads again for decode, D2H of every decoded RGB chunk; the
           chunks are split over a few host threads, each with its own batch, so copies overlap the rANS kernels.
`roofline` = the dominant kernel of the step by device time (a rANS launch); `roofline_by_kernel` lists all four
           stages, incl. the wavelet/quantise front-end and back-end, whose algorithmic traffic is 6 B per pixel
           (3 in + 3 out); achieved = algorithmic bytes / CUDA-event time, peak = measured HBM copy.
`cpu_baseline` = the C oracle (a port of the reference's single-threaded CPU path) on this box's host.
Multi-GPU: chunks are independent, so each rank runs its own batch (weak scaling), no collective on the data
path; torch.distributed (NCCL) is used only for the barrier and the max-over-ranks of the timed region.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, F = 1920, 1080, 64
QUALITY, WAVELET = 80, "cdf97"
SEED = 0x5EED0001
BYTES_PER_PX_ALG = 6.0          # 3 B in + 3 B out per RGB pixel, front-end or back-end
WORKLOAD = "CDF 9/7 q=80 1920x1080x64 encode+decode (BASELINE.json configs[1])"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU every 200 ms while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz, self.err = index, [], set(), False, None, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                     0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
                time.sleep(0.2)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "error": self.err or "no samples"}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------ CPU arms
def _oracle_chunk_seconds(frames, seed, quality=QUALITY, wavelet=1):
    """One encode+decode round trip of a 1920x1080x`frames` G1 volume through the CPU oracle; returns seconds."""
    import oracle as O
    rgb = O.generate(O.G1, W, H, frames, seed)
    t0 = time.perf_counter()
    alc = O.encode(rgb, W, H, frames, quality, wavelet)
    out = O.decode(alc)
    dt = time.perf_counter() - t0
    assert out.size == rgb.size
    return dt


def cpu_baseline_single(frames=32):
    dt = _oracle_chunk_seconds(frames, SEED)
    return {"value": round(frames / dt, 3), "unit": "frames/s", "cores": 1, "kind": "port",
            "sample": f"one 1920x1080x{frames} G1 volume (half a chunk), CDF 9/7 q=80, encode+decode through the C oracle "
                      f"(single thread, like the reference); {dt:.1f} s"}


def run_reference_arm(args):
    """--impl reference: the reference's CPU path on all host threads (the reference is a Rust crate and no Rust
    toolchain exists in this image, so this is the C oracle port; the reference is single-threaded, so the threads
    run independent chunk slices in parallel, one per worker)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import psutil
    from concurrent.futures import ThreadPoolExecutor
    frames = args.ref_frames
    per_worker_gb = 0.020 * frames * 1.1 + 0.3       # rgb + i16 planes + i32 volume + quantised + symbols (+ copies)
    avail_gb = psutil.virtual_memory().available / 2**30
    workers = max(1, min(os.cpu_count() or 1, int(avail_gb * 0.6 / per_worker_gb), 256))
    import oracle as O
    O.lib()
    times = []
    with ThreadPoolExecutor(workers) as ex:
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            list(ex.map(lambda i: _oracle_chunk_seconds(frames, SEED + i), range(workers)))  # ctypes releases the GIL
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append(dt)
    total = sum(times)
    value = workers * frames * len(times) / total
    sample = (f"{workers} worker threads x one 1920x1080x{frames} G1 slice each per step, "
              f"CDF 9/7 q=80, encode+decode through the C oracle")
    line = {"impl": "reference", "metric": "1080p encode+decode frames/s", "value": round(value, 3), "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(1000 * total / len(times), 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": sample},
            "cpu_baseline": {"value": round(value, 3), "unit": "frames/s", "cores": workers, "kind": "port",
                             "sample": sample},
            "e2e": {"value": round(value, 3), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--chunks", type=int, default=0, help="chunks in flight per GPU (0 = sized from free HBM)")
    ap.add_argument("--e2e-chunks", type=int, default=0, help="chunks per e2e step (0 = same as --chunks)")
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--e2e-threads", type=int, default=3)
    ap.add_argument("--e2e-stagger", type=float, default=2.3, help="seconds between the first starts of the e2e host threads")
    ap.add_argument("--ref-frames", type=int, default=16)
    ap.add_argument("--cpu-frames", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--in-place", action="store_true",
                    help="experiment: encode from and decode into the same RGB buffer (0.53 instead of 0.93 GB per chunk "
                         "in flight, so more rANS streams run concurrently); the inputs are regenerated on the device "
                         "inside the timed region at the start of every step")
    ap.add_argument("--max-chunks", type=int, default=0, help="cap on chunks in flight per GPU (0 = the default cap)")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--lib", default=None, help="experiment build of libalice_codec (debugging aid)")
    ap.add_argument("--e2e-only", action="store_true", help="skip the device-resident timed region (debug)")
    ap.add_argument("--quality", type=int, default=QUALITY)
    ap.add_argument("--wavelet", default=WAVELET)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3                      # timing rule: at least 3 warm-up steps
    if args.e2e_only:                        # debugging aid: the printed `value` is then not a valid measurement
        args.warmup, args.steps = 0, 1
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    pkg = load_package()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libalice_codec has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL may print its version banner on stdout when the communicator is created; stdout carries exactly one
        # JSON line, so route fd 1 to stderr until the first collective has run
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.all_reduce(torch.zeros(1, device="cuda"))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    api = pkg.Api(args.lib) if args.lib else pkg.default_api()
    api.set_device(local)
    stream = torch.cuda.current_stream()

    n_px = W * H * F
    rgb_bytes = 3 * n_px
    free_b, total_b = torch.cuda.mem_get_info()
    # per chunk in flight: RGB in + RGB out (bench buffers; the output buffer doubles as the chunk's symbol-plane
    # workspace, ALICE_BATCH_SHARED_WORKSPACE) + payload budget (1 B/px + 192 KiB) + tables
    per_chunk = (1 if args.in_place else 2) * rgb_bytes + n_px + 3 * 65536 + 3 * (16384 + 256 * 16 + 1024)
    fixed = 12 * n_px + (2 << 30)            # 4-byte scratch volume x 3 channels + head-room
    cap = args.max_chunks or (197 if args.in_place else 176)   # 197 = 592 concurrent decoder streams / 3
    B = args.chunks or max(1, min(cap, int((free_b - fixed) // per_chunk)))
    d_in = [torch.empty(rgb_bytes, dtype=torch.uint8, device="cuda") for _ in range(B)]
    d_out = d_in if args.in_place else [torch.empty(rgb_bytes, dtype=torch.uint8, device="cuda") for _ in range(B)]

    def synth_inputs():
        for i, t in enumerate(d_in):
            # chunk ids follow the round-robin sharding of alice_codec_b200.sharding: rank r holds chunks r, r+world, ...
            api._chk(api.lib.alice_codec_synth_rgb_device(1, SEED + rank + i * world, W, H, F, C.c_void_p(t.data_ptr()),
                                                          C.c_void_p(stream.cuda_stream)))
    synth_inputs()
    torch.cuda.synchronize()
    batch = pkg.ChunkBatch(args.quality, args.wavelet, W, H, F, B, stream=stream.cuda_stream, api=api,
                           shared_workspace=True)
    assert batch.workspace_bytes() == rgb_bytes
    in_ptrs = [t.data_ptr() for t in d_in]
    out_ptrs = [t.data_ptr() for t in d_out]

    def step():
        if args.in_place:
            synth_inputs()                         # the previous step decoded over the inputs
        batch.encode_device(in_ptrs, out_ptrs)     # symbol planes of chunk i live in its output buffer ...
        batch.decode_device(out_ptrs)              # ... until the decode back-end overwrites them with RGB

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ms = np.zeros(8)
    ev0.record(stream)
    for _ in range(args.steps):
        step()
        stage_ms += np.array(batch.timings())
    ev1.record(stream)
    barrier()
    sampler.stop_flag = True
    ms_total = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    sampler.join(timeout=2)
    stage_ms /= args.steps
    ms_step = ms_total / args.steps
    frames_per_step = B * F * world
    value = frames_per_step / (ms_step / 1000.0)

    # ---- correctness of what was just timed (bit-exactness is part of the metric)
    bit_exact = None
    if rank == 0:
        golden_path = os.path.join(ROOT, "tests", "golden", "fullsize.json")
        if os.path.exists(golden_path) and (args.quality, args.wavelet) == (QUALITY, WAVELET):
            import hashlib
            g = json.load(open(golden_path)).get("cfg2_cdf97_q80_1080p64")
            if g:
                alc = batch.get_chunk(0).to_bytes()
                bit_exact = (hashlib.sha256(alc).hexdigest() == g["sha256_alc"] and
                             hashlib.sha256(d_out[0].cpu().numpy().tobytes()).hexdigest() == g["sha256_decoded"])
    # alice_codec_psnr (ffi.rs:270) of chunk 0's decode against its input, computed on the device (outside the timed
    # region).  Informational: the reference's decode of its own stream is not a reconstruction and the
    # CUDA path reproduces exactly that output, so a low figure here is the reference's, not a defect.
    psnr_db = None
    if rank == 0 and not args.in_place:
        try:
            psnr_db = round(api.psnr_device(d_in[0].data_ptr(), d_out[0].data_ptr(), int(d_in[0].numel()),
                                            stream.cuda_stream), 3)
        except Exception:
            psnr_db = None

    # ---- rooflines, from the library's own CUDA events on the launch stream (averaged over the timed steps).
    # `roofline` describes the DOMINANT kernel of the step by device time (the rANS decode launch: one warp per
    # (chunk, channel) stream, a serial recurrence); `roofline_by_kernel` lists every stage, including the
    # wavelet/quantise front-end and back-end that the 6 B/px algorithmic HBM roofline is defined for.
    peak, peak_src = _peaks()
    fe_ms, be_ms = float(stage_ms[0]), float(stage_ms[5])
    enc_ms, dec_ms = float(stage_ms[2]), float(stage_ms[4])
    alg_bytes = BYTES_PER_PX_ALG * n_px * B
    n_sym = 3 * n_px * B
    n_sm = torch.cuda.get_device_properties(local).multi_processor_count
    payload_bytes = 0
    if rank == 0:
        payload_bytes = sum(batch.get_chunk(i).compressed_size for i in range(min(B, 4))) / min(B, 4) * B
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tp) and (args.quality, args.wavelet) == (QUALITY, WAVELET):
        traffic = json.load(open(tp))

    def rl(kernel, bytes_alg, ms, traffic_bytes=None, note=None):
        ach = bytes_alg / (ms / 1000.0) / 1e9 if ms > 0 else 0.0
        d = {"bound": "hbm", "kernel": kernel, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
             "frac": round(ach / peak, 4), "traffic": traffic_bytes, "peak_source": peak_src,
             "algorithmic_bytes_per_launch": bytes_alg, "ms": round(ms, 3), "share_of_step": round(ms / ms_step, 5)}
        if note:
            d["note"] = note
        return d
    by_kernel = {
        "frontend": rl("encode front-end: k_fwd_xy + k_fwd_t_quant + k_hist_zero_bin per chunk, summed over the batch",
                       alg_bytes, fe_ms, traffic.get("frontend_dram_bytes_per_chunk", 0) * B or None,
                       "6 B/px algorithmic (3 in + 3 out); latency and issue bound, not HBM bound: see DESIGN.md 4.2 and profiles/r01_pipe_stall_analysis.md"),
        "backend": rl("decode back-end: k_inv_t + k_inv_yx per chunk, summed over the batch", alg_bytes, be_ms,
                      traffic.get("backend_dram_bytes_per_chunk", 0) * B or None, "6 B/px algorithmic"),
        "rans_encode": rl("k_rans_encode (one launch, 3 streams per chunk)", n_sym + payload_bytes, enc_ms, None,
                          "serial recurrence per stream: latency bound, symbols/s per lane is the figure of merit"),
        "rans_decode": rl("k_rans_decode (one launch, 3 streams per chunk)", n_sym + payload_bytes, dec_ms, None,
                          "serial recurrence per stream: latency bound, symbols/s per lane is the figure of merit"),
    }
    dominant = max(by_kernel, key=lambda k: by_kernel[k]["ms"])
    roofline = dict(by_kernel[dominant])
    roofline["dominant_of"] = {k: v["share_of_step"] for k, v in by_kernel.items()}
    stages = {"frontend_ms": round(fe_ms, 3), "tables_enc_ms": round(float(stage_ms[1]), 3),
              "rans_encode_ms": round(enc_ms, 3), "tables_dec_ms": round(float(stage_ms[3]), 3),
              "rans_decode_ms": round(dec_ms, 3), "backend_ms": round(be_ms, 3),
              "rans_lanes": 3 * B,
              "rans_encode_msym_s_per_lane": round(n_px / (enc_ms / 1000.0) / 1e6, 2) if enc_ms > 0 else None,
              "rans_decode_msym_s_per_lane": round(n_px / (dec_ms / 1000.0) / 1e6, 2) if dec_ms > 0 else None,
              "rans_encode_msym_s_per_sm": round(n_sym / (enc_ms / 1000.0) / 1e6 / n_sm, 2) if enc_ms > 0 else None,
              "rans_decode_msym_s_per_sm": round(n_sym / (dec_ms / 1000.0) / 1e6 / n_sm, 2) if dec_ms > 0 else None}

    # ---- e2e: the same metric through the host-buffer C-ABI call, copies inside the timed region
    e2e = None
    if args.in_place and not args.no_e2e:
        synth_inputs()                                  # the last step decoded over the inputs the e2e leg copies out
        torch.cuda.synchronize()
    if not args.no_e2e:
        # The caller-side pattern for host buffers: T host threads, each driving its own ChunkBatch (own CUDA stream)
        # over its share of the chunks, so one batch's PCIe copies run under another batch's rANS kernels.
        Be = args.e2e_chunks or B
        try:                                            # pinned RGB in + out must fit the host comfortably (all ranks)
            import psutil
            host_avail = psutil.virtual_memory().available
            Be = max(1, min(Be, int(host_avail * (0.6 if world == 1 else 0.5) / world // (2 * rgb_bytes))))
        except Exception:  # noqa: BLE001
            Be = min(Be, 32)
        T = max(1, min(args.e2e_threads, Be))
        h_in = [torch.empty(rgb_bytes, dtype=torch.uint8).pin_memory() for _ in range(Be)]
        h_out = [torch.empty(rgb_bytes, dtype=torch.uint8).pin_memory() for _ in range(Be)]
        for i in range(Be):
            h_in[i].copy_(d_in[i])
        torch.cuda.synchronize()
        # link probe: one pinned chunk each way, three times (explains run-to-run differences of the e2e figure)
        link = {}
        for name, fn in (("h2d", lambda: d_in[0].copy_(h_in[0], non_blocking=True)),
                         ("d2h", lambda: h_out[0].copy_(d_out[0], non_blocking=True))):
            fn()
            torch.cuda.synchronize()
            tp0 = time.perf_counter()
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            link[name + "_gbs"] = round(3 * rgb_bytes / (time.perf_counter() - tp0) / 1e9, 1)
        h_in[0].copy_(d_in[0])                          # restore nothing: d_in[0] was overwritten with its own data
        torch.cuda.synchronize()
        batch.close()
        del d_in, d_out, in_ptrs, out_ptrs
        torch.cuda.empty_cache()
        shares = [list(range(t, Be, T)) for t in range(T)]
        streams = [torch.cuda.Stream() for _ in range(T)]
        batches = [pkg.ChunkBatch(args.quality, args.wavelet, W, H, F, len(shares[t]), stream=streams[t].cuda_stream,
                                  api=api, shared_workspace=True) for t in range(T)]
        payload = [0] * T
        t_start, t_end = [0.0] * T, [0.0] * T
        go = threading.Barrier(T + 1)

        def worker(t):
            api.set_device(local)                       # cudaSetDevice is per host thread
            hin = [h_in[i].data_ptr() for i in shares[t]]
            hout = [h_out[i].data_ptr() for i in shares[t]]
            go.wait()
            time.sleep(t * args.e2e_stagger)            # start the batches out of phase: copies of one run under
            for it in range(1 + args.e2e_steps):        # the rANS kernels of the others.  Iteration 0 = warm-up.
                if it == 1:
                    t_start[t] = time.perf_counter()
                ta = time.perf_counter()
                chunks = batches[t].encode_host(hin)    # H2D RGB, kernels, D2H headers + payload
                tb = time.perf_counter()
                payload[t] = sum(c.compressed_size + 3138 for c in chunks)
                batches[t].decode_host(chunks, hout)    # H2D payload, kernels, D2H RGB
                del chunks                              # payload buffers go back to the library's pinned pool
                tc = time.perf_counter()
                if args.verbose:
                    ms = batches[t].timings()
                    sys.stderr.write(f"[e2e] thread {t} it {it}: encode_host {tb - ta:.2f} s (fe {ms[0] / 1e3:.2f} rans {ms[2] / 1e3:.2f}), "
                                     f"decode_host {tc - tb:.2f} s (rans {ms[4] / 1e3:.2f} be {ms[5] / 1e3:.2f})\n")
            t_end[t] = time.perf_counter()
        threads = [threading.Thread(target=worker, args=(t,)) for t in range(T)]
        for th in threads:
            th.start()
        barrier()
        go.wait()
        for th in threads:
            th.join()
        barrier()
        # the timed window runs from the first thread's first timed iteration to the last thread's last one; parts
        # of other threads' warm-up iterations that fall inside it are not counted as work
        dt = max(t_end) - min(t_start)
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e_ok = None
        if rank == 0 and bit_exact is not None:
            import hashlib
            e2e_ok = hashlib.sha256(h_out[0].numpy().tobytes()).hexdigest() == g["sha256_decoded"]
        e2e = {"value": round(Be * F * world * args.e2e_steps / dt, 2), "unit": "frames/s",
               "h2d_bytes_per_step": Be * rgb_bytes + sum(payload), "d2h_bytes_per_step": Be * rgb_bytes + sum(payload),
               "chunks_per_step_per_gpu": Be, "host_threads": T, "ms_per_step": round(1000 * dt / args.e2e_steps, 1),
               "timed_with": "host wall clock around synchronous C-ABI batch calls (pinned host buffers)",
               "decoded_matches_oracle_digest": e2e_ok, "pinned_link_probe": link}
        for bt in batches:
            bt.close()
        del h_in, h_out
        batch = None

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_single(args.cpu_frames)

    if rank == 0:
        line = {"metric": "1080p encode+decode frames/s", "value": round(value, 2), "unit": "frames/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 3),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "chunks_per_gpu_per_step": B, "frames_per_step": frames_per_step,
                           "input": "G1 tri+hash volumes generated on the device, one seed per chunk",
                           "l2": "inputs (%.1f GB per step per GPU) are far larger than the 126 MB L2" % (B * rgb_bytes / 1e9),
                           "parallelism": f"chunks sharded over {world} GPU(s), no data-path collective",
                           **({"in_place": "inputs regenerated on the device inside every timed step"} if args.in_place else {})},
                "e2e": e2e, "gpu_launches": args.steps * ((6 if args.in_place else 5) * B + 4), "roofline": roofline, "roofline_by_kernel": by_kernel,
                "stages": stages,
                "cpu_baseline": cpu, "clocks": sampler.summary(), "bit_exact_vs_oracle_digest": bit_exact,
                "psnr_decoded_vs_input_db": psnr_db}
        print(json.dumps(line), flush=True)
    if batch is not None:
        batch.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
