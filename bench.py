#!/usr/bin/env python
"""bench.py -- TERSE (encode) + PROLIX (decode) throughput on B200, BASELINE.json's metric:
frames/s and uncompressed GB/s, 512x512 uint16 stack, device-resident and PCIe-inclusive, with the
HBM roofline of the dominant kernel and the reference CPU codec timed on the same box.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--frames F]      # ours (CUDA, C ABI)
  python bench.py --impl reference [...]                               # the reference's CPU path

A "step" is one pass of the hot path over one batch: TERSE-encode all F frames, then PROLIX-decode
them all.  N > 1 (torchrun, one process per GPU): every rank owns its own F frames (frames are
independent: no collective on the data path; "scaling": "weak"), time = max over ranks.

Inputs are synthetic (Poisson(2) background + 200 Gaussian Bragg peaks per frame), generated on the
device with torch before the timed region.  torch is plumbing only (memory, events, distributed);
every timed kernel is launched by libtrpx_b200.so."""
import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 512, 512
N_VALUES = H * W
LAMBDA, N_PEAKS = 2.0, 200
METRIC = "terse+prolix frames/s, 512x512 uint16 stack (encode then decode every frame)"
UNIT = "frames/s"
WORKLOAD = "configs[1]: 10,000-frame 512x512 uint16 stack, compress+decompress on 1 B200 (per GPU)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=10000, help="frames per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-chunk", type=int, default=1000,
                    help="frames per chunk of the streamed e2e (encode of chunk k+1 overlaps decode of chunk k)")
    ap.add_argument("--e2e-ramp", type=int, default=0, help="frames of the first and of the last chunk (0: --e2e-chunk)")
    ap.add_argument("--e2e-enc-threads", type=int, default=1)
    ap.add_argument("--e2e-dec-threads", type=int, default=1)
    ap.add_argument("--cpu-sample", type=int, default=3000, help="frames of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ inputs
def synth_stack(torch, frames, seed, dev, chunk=200):
    """(frames, 512*512) int16 tensor holding the uint16 bit patterns of Poisson + Bragg-peak frames."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    out = torch.empty((frames, N_VALUES), dtype=torch.int16, device=dev)
    r = torch.arange(-4, 5, device=dev, dtype=torch.float32)
    dy, dx = torch.meshgrid(r, r, indexing="ij")
    dy, dx = dy.reshape(1, 1, -1), dx.reshape(1, 1, -1)
    for f0 in range(0, frames, chunk):
        n = min(chunk, frames - f0)
        img = torch.poisson(torch.full((n, N_VALUES), LAMBDA, device=dev), generator=g)
        cy = torch.rand((n, N_PEAKS, 1), device=dev, generator=g) * (H - 1)
        cx = torch.rand((n, N_PEAKS, 1), device=dev, generator=g) * (W - 1)
        sigma = 1.0 + torch.rand((n, N_PEAKS, 1), device=dev, generator=g)
        amp = torch.exp(math.log(20.0) + torch.rand((n, N_PEAKS, 1), device=dev, generator=g) * math.log(3000.0 / 20.0))
        iy, ix = cy.round() + dy, cx.round() + dx
        val = amp * torch.exp(-((iy - cy) ** 2 + (ix - cx) ** 2) / (2 * sigma * sigma))
        val = torch.poisson(val, generator=g)
        ok = (iy >= 0) & (iy < H) & (ix >= 0) & (ix < W)
        idx = (iy.clamp(0, H - 1) * W + ix.clamp(0, W - 1)).long().reshape(n, -1)
        img.scatter_add_(1, idx, (val * ok).reshape(n, -1))
        out[f0:f0 + n] = img.clamp_(0, 65535).to(torch.int32).to(torch.int16)   # int -> int wraps: u16 bit pattern
    return out


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock, power and throttle reasons of one GPU sampled DURING the timed region: an in-process NVML thread
    (a sample every ~2 ms, time-stamped), or -- without pynvml -- an `nvidia-smi -lms 20` child process.  Started
    well before the timed region (NVML / nvidia-smi need up to a second to come up on an 8-GPU box); begin() and
    stop() bracket the region and only samples taken in between are reported."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, uuid=None):
        self.idx = gpu_index
        self.uuid = uuid
        self.proc = None
        self.nvml = None
        self.samples = []            # (t, sm_mhz, sm_max_mhz, power_w, reasons tuple)
        self.t_begin = 0.0
        self.run = False

    def start(self):
        try:
            import pynvml as N
            N.nvmlInit()
            h = None
            if self.uuid:
                for u in ("GPU-" + str(self.uuid), str(self.uuid)):
                    try:
                        h = N.nvmlDeviceGetHandleByUUID(u.encode() if isinstance(u, str) else u)
                        break
                    except Exception:
                        h = None
            if h is None:
                h = N.nvmlDeviceGetHandleByIndex(self.idx)
            smax = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
            get_reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

            def loop():
                while self.run:
                    try:
                        r = int(get_reasons(h))
                        self.samples.append((time.perf_counter(), float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)), smax,
                                             N.nvmlDeviceGetPowerUsage(h) / 1000.0, tuple(n for n, b in bits if r & b)))
                    except Exception:
                        pass
                    time.sleep(0.002)

            self.nvml = N
            self.run = True
            self.t = threading.Thread(target=loop, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                self.samples.append((time.perf_counter(), float(f[1]), float(f[2]), float(f[3]),
                                     tuple(n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                                              "sw_power_cap"), f[5:9]) if v.lower().startswith("active"))))
            except ValueError:
                continue

    def begin(self):
        self.t_begin = time.perf_counter()

    def stop(self):
        t_end = time.perf_counter()
        if self.nvml is not None:
            self.run = False
            self.t.join(timeout=1)
            how = "NVML thread"
        elif self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
            how = "nvidia-smi -lms 20"
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML and nvidia-smi unavailable"], "samples": 0}
        inside = [x for x in self.samples if self.t_begin <= x[0] <= t_end + (0.0 if self.nvml is not None else 0.05)]
        window = "timed region"
        if not inside:                                          # (nvidia-smi only: region shorter than its period)
            inside = [x for x in self.samples if x[0] >= self.t_begin - 0.1][:3] or self.samples[-3:]
            window = "nearest samples"
        sm = sorted(x[1] for x in inside)
        reasons = sorted({r for x in inside for r in x[4]})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(x[2] for x in inside) if inside else None,
                "reasons": reasons, "samples": len(inside), "power_w_max": max(x[3] for x in inside) if inside else None,
                "window": window, "how": how}


# ------------------------------------------------------------------------------------------ CPU baseline
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_codec_time(px_np, threads, repeats=2):
    """Reference CPU codec (oracle/_ref: the reference's own headers, -O3 -DNDEBUG; one jpa::Terse per
    frame, frames statically partitioned over `threads` std::threads) on px_np (F, N) uint16.
    Falls back to the C port (oracle/liboracle.so) when _ref was not built.
    -> (kind, enc_s, dec_s, payload_bytes)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import orc                                    # TEST-ONLY checker: used here as the timed CPU baseline
    F, N = px_np.shape
    ref = orc.ref()
    if ref is not None:
        best_e = best_d = float("inf")
        tot = ctypes.c_size_t(0)
        out = np.empty_like(px_np)
        for _ in range(repeats):
            best_e = min(best_e, ref.ref_bench_encode(px_np.ctypes.data, orc.U16, N, F, threads, ctypes.byref(tot)))
            best_d = min(best_d, ref.ref_bench_decode(px_np.ctypes.data, orc.U16, N, F, threads, out.ctypes.data))
        assert np.array_equal(out, px_np), "reference CPU round trip failed"
        return "reference", best_e, best_d, int(tot.value)
    from concurrent.futures import ThreadPoolExecutor
    orc.build()
    payloads = [None] * F

    def enc(f):
        payloads[f] = orc.encode_frame(px_np[f])[0]

    def dec(f):
        orc.decode_frame(payloads[f], N, False, np.uint16)

    with ThreadPoolExecutor(threads) as ex:
        t0 = time.perf_counter()
        list(ex.map(enc, range(F)))
        t1 = time.perf_counter()
        list(ex.map(dec, range(F)))
        t2 = time.perf_counter()
    return "port", t1 - t0, t2 - t1, sum(p.size for p in payloads)


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def synth_stack_cpu(frames, seed):
    """CPU-side frames of the same distribution for the reference arm (no GPU needed)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    img = rng.poisson(LAMBDA, size=(frames, H, W)).astype(np.float32)
    yy, xx = np.mgrid[-4:5, -4:5]
    for f in range(frames):
        cy, cx = rng.random(N_PEAKS) * (H - 1), rng.random(N_PEAKS) * (W - 1)
        sg = 1.0 + rng.random(N_PEAKS)
        amp = np.exp(math.log(20.0) + rng.random(N_PEAKS) * math.log(150.0))
        for k in range(N_PEAKS):
            iy, ix = int(round(cy[k])) + yy, int(round(cx[k])) + xx
            ok = (iy >= 0) & (iy < H) & (ix >= 0) & (ix < W)
            v = rng.poisson(amp[k] * np.exp(-((iy - cy[k]) ** 2 + (ix - cx[k]) ** 2) / (2 * sg[k] ** 2)))
            np.add.at(img[f], (iy[ok], ix[ok]), v[ok])
    return np.clip(img, 0, 65535).astype(np.uint16).reshape(frames, N_VALUES)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    sample = max(cores, min(a.frames, 1500))
    px = synth_stack_cpu(min(sample, 64), 4242)
    import numpy as np
    px = np.ascontiguousarray(np.tile(px, (int(math.ceil(sample / px.shape[0])), 1))[:sample])
    times = []
    kind = "port"
    for i in range(a.warmup + a.steps):
        kind, te, td, cbytes = cpu_codec_time(px, cores, repeats=1)
        if i >= a.warmup:
            times.append((te, td))
    te = sum(t[0] for t in times) / len(times)
    td = sum(t[1] for t in times) / len(times)
    v = sample / (te + td)
    desc = "%d frames (64 distinct synthetic frames tiled) per step, encode then decode, %d threads" % (sample, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * (te + td), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frame": "512x512", "pixel": "uint16", "block": 12, "sample": desc},
        "encode_frames_per_s": sample / te, "decode_frames_per_s": sample / td,
        "uncompressed_GBps": v * N_VALUES * 2 / 1e9, "compression_ratio": cbytes / (sample * N_VALUES * 2.0),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc, "cpu": cpu_model()},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------ ours
def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def run_ours(a):
    import numpy as np
    import torch
    import trpx_b200
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this codec has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    F = a.frames

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    try:
        gpu_uuid = torch.cuda.get_device_properties(local).uuid
    except Exception:
        gpu_uuid = None
    sampler = ClockSampler(local, gpu_uuid)
    sampler.start()                                                 # begin() marks the timed region
    codec = trpx_b200.Codec(local)
    codec.set_profiling(True)
    px = synth_stack(torch, F, 1000 + 100000 * rank, dev)
    cap = trpx_b200.max_compressed_bytes(N_VALUES, np.uint16, 12, F)
    payload = torch.empty(cap, dtype=torch.uint8, device=dev)
    ends = torch.zeros(F, dtype=torch.int64, device=dev)
    small = torch.zeros(4, dtype=torch.int32, device=dev)          # prolix_bits, enc status, dec status
    back = torch.empty((F, N_VALUES), dtype=torch.int16, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    raw_bytes = F * N_VALUES * 2

    def encode():
        codec.encode_device(px.data_ptr(), np.uint16, N_VALUES, F, payload.data_ptr(), cap, ends.data_ptr(),
                            small.data_ptr(), small.data_ptr() + 4, stream)

    def decode(nbytes):
        codec.decode_device(payload.data_ptr(), nbytes, False, N_VALUES, F, ends.data_ptr(), back.data_ptr(),
                            np.uint16, small.data_ptr() + 8, stream, lane=1)

    # ---- untimed: first pass, correctness of the full-size workload (round trip on the device)
    encode()
    torch.cuda.synchronize()
    cbytes = int(ends[F - 1])
    assert int(small[1]) == 0, "encode status %d" % int(small[1])
    decode(cbytes)
    torch.cuda.synchronize()
    assert int(small[2]) == 0, "decode status %d" % int(small[2])
    assert torch.equal(back, px), "PROLIX(TERSE(x)) != x"
    back.zero_()
    for _ in range(max(a.warmup, 3) - 1):
        encode()
        decode(cbytes)
    barrier()

    # ---- timed: exactly K steps, CUDA events on the launching stream
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(a.steps)]
    ktimes = {}
    sampler.begin()
    l0 = codec.launches
    t_host0 = time.perf_counter()
    for k in range(a.steps):
        ev[k][0].record()
        encode()
        ev[k][1].record()
        decode(cbytes)
        ev[k][2].record()
        if a.steps <= 64:
            # per-kernel device times come from events the library drops between its kernels; reading
            # them needs a drained stream, so this costs one host sync per step (outside the kernels)
            torch.cuda.synchronize()
            for name, ms in codec.last_kernel_times(0) + codec.last_kernel_times(1):
                ktimes.setdefault(name, []).append(ms)
    torch.cuda.synchronize()
    launches = codec.launches - l0
    clocks = sampler.stop()
    t_host1 = time.perf_counter()
    barrier()
    enc_ms = [ev[k][0].elapsed_time(ev[k][1]) for k in range(a.steps)]
    dec_ms = [ev[k][1].elapsed_time(ev[k][2]) for k in range(a.steps)]
    step_ms = [e + d for e, d in zip(enc_ms, dec_ms)]
    total_ms = sum(step_ms)
    assert int(small[1]) == 0 and int(small[2]) == 0
    assert torch.equal(back, px), "round trip failed after the timed steps"
    tmax = torch.tensor([total_ms, sum(enc_ms), sum(dec_ms)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms_max, enc_ms_max, dec_ms_max = [float(x) for x in tmax.cpu()]

    # ---- e2e: host buffers (pinned), through the host-pointer C ABI; H2D + D2H inside the timed region
    e2e = None
    if not a.no_e2e:
        # the whole stack in pinned host memory (11.9 GB per rank at 10,000 frames); should the host refuse that much
        # (all ranks of a box pin at once), the e2e runs on the first half, quarter ... of the frames and says so
        Fe = F
        while True:
            try:
                cb_e = int(ends[Fe - 1])
                pfs = (int(1.25 * cb_e / Fe) + 255) // 256 * 256   # payload slot bytes per frame of the streamed run
                h_px = torch.empty((Fe, N_VALUES), dtype=torch.int16, pin_memory=True)
                h_back = torch.empty((Fe, N_VALUES), dtype=torch.int16, pin_memory=True)
                h_payload = torch.empty(Fe * pfs + (Fe + 2) * 4096, dtype=torch.uint8, pin_memory=True)
                break
            except RuntimeError:
                h_px = h_back = h_payload = None
                if Fe <= 500:
                    raise
                Fe //= 2
        if dist is not None:                                       # every rank runs the e2e on the same number of frames
            fe_t = torch.tensor([Fe], dtype=torch.int64, device=dev)
            dist.all_reduce(fe_t, op=dist.ReduceOp.MIN)
            if int(fe_t[0]) < Fe:
                Fe = int(fe_t[0])
                cb_e = int(ends[Fe - 1])
                h_px, h_back = h_px[:Fe], h_back[:Fe]
        raw_e = Fe * N_VALUES * 2
        h_px.copy_(px[:Fe])
        fb = np.zeros(Fe, np.uint64)
        L = trpx_b200.lib()
        tot = ctypes.c_size_t(0)
        pb = ctypes.c_uint(0)
        # device staging of the timed kernels above is no longer needed
        frame_raw = N_VALUES * 2

        seq_parts = []

        def sequential():
            """encode the whole stack, then decode the whole payload: two calls on one context"""
            t_a = time.perf_counter()
            rc = L.trpx_encode_host(codec._h, h_px.data_ptr(), trpx_b200.U16, N_VALUES, Fe, 12, h_payload.data_ptr(),
                                    h_payload.numel(), fb.ctypes.data, ctypes.byref(tot), ctypes.byref(pb))
            assert rc == 0, rc
            t_b = time.perf_counter()
            rc = L.trpx_decode_host(codec._h, h_payload.data_ptr(), tot.value, 0, 12, N_VALUES, Fe, 0, Fe,
                                    fb.ctypes.data, None, h_back.data_ptr(), trpx_b200.U16)
            assert rc == 0, rc
            seq_parts.append((t_b - t_a, time.perf_counter() - t_b))
            return tot.value

        # Streamed: the stack goes through in chunks of --e2e-chunk frames.  Encoder threads (one context each, "one
        # context per host thread", trpx_b200.h) take the chunks round-robin; every chunk's payload lands in its own
        # slot of the pinned payload buffer; decoder threads (again one context each) decode chunk k as soon as it is
        # encoded.  The pixel H2D of the encoders and the pixel D2H of the decoders then share the full-duplex PCIe
        # link instead of taking turns, and one call's fill/drain is covered by its neighbour's copies.
        n_enc, n_dec = max(1, a.e2e_enc_threads), max(1, a.e2e_dec_threads)
        enc_ctx = [codec] + [trpx_b200.Codec(local) for _ in range(n_enc - 1)]
        dec_ctx = [trpx_b200.Codec(local) for _ in range(n_dec)]
        chunk = max(1, min(Fe, a.e2e_chunk))
        ramp = max(1, min(chunk, a.e2e_ramp or chunk))             # (a shorter first and last chunk did not pay: DESIGN.md 5)
        cuts = [0] + list(range(ramp, Fe - ramp, chunk)) + ([Fe - ramp] if Fe > 2 * ramp else []) + [Fe]
        cuts = sorted(set(cuts))
        n_chunks = len(cuts) - 1
        slot_off = [cuts[c] * pfs + c * 4096 for c in range(n_chunks)]
        h_slots = h_payload                                        # (sized for the slots below)
        assert h_slots.numel() >= Fe * pfs + (n_chunks + 1) * 4096

        trace = []

        def streamed():
            import queue
            qs = [queue.Queue() for _ in range(n_dec)]
            err = []
            sizes = [0] * n_chunks
            del trace[:]
            t_ref = time.perf_counter()

            def enc_side(e):
                t_, p_ = ctypes.c_size_t(0), ctypes.c_uint(0)
                for c in range(e, n_chunks, n_enc):
                    f0, nf = cuts[c], cuts[c + 1] - cuts[c]
                    t_s = time.perf_counter()
                    rc = L.trpx_encode_host(enc_ctx[e]._h, h_px.data_ptr() + f0 * frame_raw, trpx_b200.U16, N_VALUES, nf,
                                            12, h_slots.data_ptr() + slot_off[c], nf * pfs + 4096, fb.ctypes.data + 8 * f0,
                                            ctypes.byref(t_), ctypes.byref(p_))
                    if rc != 0:
                        err.append(("encode", c, rc))
                    sizes[c] = t_.value
                    trace.append(("enc", c, round(1e3 * (t_s - t_ref), 2), round(1e3 * (time.perf_counter() - t_ref), 2)))
                    qs[c % n_dec].put((c, f0, nf, t_.value if rc == 0 else 0))

            def dec_side(d):
                for _ in range(d, n_chunks, n_dec):
                    c, f0, nf, nbytes = qs[d].get()
                    if not nbytes:
                        continue
                    t_s = time.perf_counter()
                    rc = L.trpx_decode_host(dec_ctx[d]._h, h_slots.data_ptr() + slot_off[c], nbytes, 0, 12, N_VALUES, nf, 0,
                                            nf, fb.ctypes.data + 8 * f0, None, h_back.data_ptr() + f0 * frame_raw,
                                            trpx_b200.U16)
                    trace.append(("dec", c, round(1e3 * (t_s - t_ref), 2), round(1e3 * (time.perf_counter() - t_ref), 2)))
                    if rc != 0:
                        err.append(("decode", c, rc))

            th = [threading.Thread(target=enc_side, args=(e,)) for e in range(n_enc)] + \
                 [threading.Thread(target=dec_side, args=(d,)) for d in range(n_dec)]
            for t_ in th:
                t_.start()
            for t_ in th:
                t_.join()
            assert not err, err
            return sum(sizes)

        # Overlapped: ONE trpx_encode_host call for the whole stack (a single uninterrupted upload pipeline); a second
        # host thread follows trpx_ctx_encode_progress() and decodes, on its own context, whatever prefix of the stack
        # has landed since its last call (at least --e2e-chunk frames unless the encoder has finished).
        def overlapped():
            err = []
            seq0 = codec.encode_progress()[0]
            enc_done = threading.Event()

            def enc_side():
                t_, p_ = ctypes.c_size_t(0), ctypes.c_uint(0)
                rc = L.trpx_encode_host(codec._h, h_px.data_ptr(), trpx_b200.U16, N_VALUES, Fe, 12, h_payload.data_ptr(),
                                        h_payload.numel(), fb.ctypes.data, ctypes.byref(t_), ctypes.byref(p_))
                if rc != 0:
                    err.append(("encode", rc))
                tot.value = t_.value
                enc_done.set()

            def dec_side():
                f_prev = b_prev = 0
                while f_prev < Fe and not err:
                    finished = enc_done.is_set()
                    seq, f_now, b_now = codec.encode_progress()
                    if seq == seq0:
                        f_now = b_now = 0                          # the encode call has not started yet
                    if f_now - f_prev >= chunk or (finished and f_now > f_prev):
                        nf = f_now - f_prev
                        rc = L.trpx_decode_host(dec_ctx[0]._h, h_payload.data_ptr() + b_prev, b_now - b_prev, 0, 12,
                                                N_VALUES, nf, 0, nf, fb.ctypes.data + 8 * f_prev, None,
                                                h_back.data_ptr() + f_prev * frame_raw, trpx_b200.U16)
                        if rc != 0:
                            err.append(("decode", f_prev, rc))
                        f_prev, b_prev = f_now, b_now
                    elif finished and f_now == f_prev:
                        err.append(("stalled", f_prev))
                    else:
                        time.sleep(0.0002)

            th = [threading.Thread(target=enc_side), threading.Thread(target=dec_side)]
            for t_ in th:
                t_.start()
            for t_ in th:
                t_.join()
            assert not err, err
            return tot.value

        def timed(fn):
            """per-step wall times of 1 warm-up + --e2e-steps timed passes (max over ranks); the median is reported:
            PCIe throughput on a shared host varies from pass to pass, all passes are listed in the JSON"""
            ts = []
            for k in range(1 + a.e2e_steps):
                h_back.zero_()
                barrier()
                t0 = time.perf_counter()
                nbytes = fn()
                torch.cuda.synchronize()
                t1 = time.perf_counter()
                assert nbytes == cb_e and torch.equal(h_back, h_px), "e2e round trip failed"
                if k:                                              # first pass = warm-up (allocations)
                    ts.append(t1 - t0)
            t = torch.tensor(ts, dtype=torch.float64, device=dev)
            if dist is not None:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts = sorted(float(x) for x in t.cpu())
            return ts[len(ts) // 2] if len(ts) % 2 else 0.5 * (ts[len(ts) // 2 - 1] + ts[len(ts) // 2]), ts

        seq_s, seq_all = timed(sequential)
        str_s, str_all = timed(streamed)
        ovl_s, ovl_all = timed(overlapped)
        for c_ in enc_ctx[1:] + dec_ctx:
            c_.close()
        if os.environ.get("TRPX_E2E_TRACE"):
            sys.stderr.write("e2e trace (side, chunk, start ms, end ms): %s\n" % sorted(trace, key=lambda r: r[2]))
        e2e_s = min(seq_s, str_s, ovl_s)
        e2e = {"value": world * Fe / e2e_s, "unit": UNIT, "h2d_bytes_per_step": raw_e + cb_e + 8 * Fe,
               "d2h_bytes_per_step": cb_e + raw_e + 16 * Fe, "ms_per_step": 1e3 * e2e_s,
               "uncompressed_GBps": world * raw_e / e2e_s / 1e9,
               "frames_per_gpu": Fe, "api": "trpx_encode_host + trpx_decode_host (pinned host buffers)",
               "mode": "overlapped" if ovl_s <= min(seq_s, str_s) else "streamed" if str_s <= seq_s else "sequential",
               "overlapped": {"value": world * Fe / ovl_s, "ms_per_step": 1e3 * ovl_s,
                              "ms_all_steps": [round(1e3 * x, 2) for x in ovl_all], "min_decode_frames": chunk,
                              "how": "one trpx_encode_host call for the whole stack; a second host thread follows "
                                     "trpx_ctx_encode_progress() and decodes the finished prefix on its own context"},
               "streamed": {"value": world * Fe / str_s, "ms_per_step": 1e3 * str_s, "ms_all_steps": [round(1e3 * x, 2) for x in str_all],
                            "chunk_frames": chunk, "first_last_chunk_frames": ramp,
                            "encoder_threads": n_enc, "decoder_threads": n_dec,
                            "how": "host threads with one context each: chunks are encoded round-robin and decoded as "
                                   "soon as they are encoded, so H2D and D2H overlap (full-duplex PCIe)"},
               "timing": "median of %d timed passes after one warm-up pass, host clock around the calls" % a.e2e_steps,
               "sequential": {"value": world * Fe / seq_s, "ms_per_step": 1e3 * seq_s,
                              "ms_all_steps": [round(1e3 * x, 2) for x in seq_all],
                              "encode_ms": 1e3 * min(p[0] for p in seq_parts[1:]),
                              "decode_ms": 1e3 * min(p[1] for p in seq_parts[1:]),
                              "how": "one call encodes the whole stack, a second one decodes the whole payload"}}
        del h_px, h_back, h_payload

    # ---- CPU baseline on this box's cores (rank 0, N == 1), and a cross-check of the payload size
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cores = host_cores()
        sample = min(F, a.cpu_sample)
        px_np = px[:sample].cpu().numpy().view(np.uint16)
        kind, te, td, cb = cpu_codec_time(px_np, cores)
        assert cb == int(ends[sample - 1]), "CPU reference payload size differs from the GPU's"
        cpu = {"value": sample / (te + td), "unit": UNIT, "cores": cores, "kind": kind,
               "sample": "first %d of the %d frames, encode then decode, best of 2" % (sample, F),
               "encode_frames_per_s": sample / te, "decode_frames_per_s": sample / td, "cpu": cpu_model()}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0
    peak, peak_src = peaks()
    alg_bytes = raw_bytes + cbytes                                  # per pass: N*sizeof(T) + C
    kavg = {n: sum(v) / len(v) for n, v in ktimes.items()}
    dom = max(kavg, key=kavg.get) if kavg else None
    roof = None
    traffic = {}
    try:                                                            # DRAM bytes per launch from the committed ncu capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        if tj.get("frames") == F:
            traffic = {k: v["traffic_bytes"] for k, v in tj["kernels"].items()}
    except Exception:
        pass
    if dom:
        ach = alg_bytes / (kavg[dom] * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic.get(dom), "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                "avg_launch_ms": kavg[dom], "frac_of_8TBps_nominal": ach / 8000.0,
                "traffic_source": "profiles/r01_traffic.json (ncu --set full, same workload)" if dom in traffic else None}
    enc_avg, dec_avg = enc_ms_max / a.steps, dec_ms_max / a.steps
    value = world * F * a.steps / (total_ms_max * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": total_ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_gpu": F, "frame": "512x512", "pixel": "uint16", "block": 12,
                   "background": "Poisson(2.0)", "bragg_peaks_per_frame": N_PEAKS,
                   "l2": "inputs larger than L2 (%.2f GB of pixels per pass vs 126 MB)" % (raw_bytes / 1e9)},
        "encode_frames_per_s": world * F / (enc_avg * 1e-3), "decode_frames_per_s": world * F / (dec_avg * 1e-3),
        "encode_uncompressed_GBps": world * raw_bytes / (enc_avg * 1e-3) / 1e9,
        "decode_uncompressed_GBps": world * raw_bytes / (dec_avg * 1e-3) / 1e9,
        "uncompressed_GBps": value * N_VALUES * 2 / 1e9,
        "compression_ratio": cbytes / float(raw_bytes), "prolix_bits": int(small[0]),
        "passes": {"encode": {"ms": enc_avg, "hbm_GBps": alg_bytes / (enc_avg * 1e-3) / 1e9,
                              "frac_of_measured_peak": alg_bytes / (enc_avg * 1e-3) / 1e9 / peak},
                   "decode": {"ms": dec_avg, "hbm_GBps": alg_bytes / (dec_avg * 1e-3) / 1e9,
                              "frac_of_measured_peak": alg_bytes / (dec_avg * 1e-3) / 1e9 / peak,
                              "traffic": (traffic.get("prolix_walk", 0) + traffic.get("prolix_unpack_seg", 0)) or None}},
        "kernel_ms": kavg, "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": clocks, "vs_readme_claim_2000_frames_per_s": value / 2000.0,
        "host_wall_s_timed_region": t_host1 - t_host0,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)
    return run_ours(a)


if __name__ == "__main__":
    sys.exit(main())
