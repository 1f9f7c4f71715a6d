#!/usr/bin/env python
"""bench.py -- TERSE (encode) + PROLIX (decode) throughput on B200, BASELINE.json's metric:
frames/s and uncompressed GB/s, device-resident and PCIe-inclusive, with the HBM roofline of both passes and the
reference CPU codec timed on the same box.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2|c3|c4u8|c4u16|c5i16|c5i32] [--frames F]
  python bench.py --impl reference [...]                               # the reference's CPU path, same config

--config c2 (default) is BASELINE.json configs[1], the configuration the metric is quoted on: 10,000 x 512x512 uint16,
encode then decode, per GPU (N > 1: every rank owns its own stack, "weak").  c3 / c4* / c5* are configs[2..4]: ONE
stack, frame-sharded over the N GPUs ("strong": contiguous frame ranges, no collective on the data path); c5* time the
decoder only (configs[4] is a decode-only sweep).

A "step" is one pass of the hot path over the batch: TERSE-encode all frames, then PROLIX-decode them all.  Inputs are
synthetic, generated on the device with torch before the timed region.  torch is plumbing only (memory, events,
distributed); every timed kernel is launched by libtrpx_b200.so."""
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

UNIT = "frames/s"
# name -> (BASELINE.json config, width, height, numpy dtype name, frames (c2: per GPU; others: the whole stack),
#          scaling, decode only, cpu_baseline sample frames, generator)
CONFIGS = {
    "c2": dict(label="configs[1]: 10,000-frame 512x512 uint16 stack, compress+decompress on 1 B200 (per GPU)",
               w=512, h=512, dtype="uint16", frames=10000, scaling="weak", decode_only=False, cpu_sample=3000, gen="bragg"),
    "c3": dict(label="configs[2]: Eiger2-16M-class 4148x4362 uint32 frames, one stack frame-sharded over the GPUs",
               w=4148, h=4362, dtype="uint32", frames=64, scaling="strong", decode_only=False, cpu_sample=8, gen="eiger"),
    "c4u8": dict(label="configs[3]: cryo-EM counting-camera 5760x4092 uint8 movie (mostly 0/1), one stack frame-sharded over the GPUs",
                 w=5760, h=4092, dtype="uint8", frames=320, scaling="strong", decode_only=False, cpu_sample=16, gen="sparse"),
    "c4u16": dict(label="configs[3]: cryo-EM counting-camera 5760x4092 uint16 movie (mostly 0/1), one stack frame-sharded over the GPUs",
                  w=5760, h=4092, dtype="uint16", frames=320, scaling="strong", decode_only=False, cpu_sample=16, gen="sparse"),
    "c5i16": dict(label="configs[4]: signed int16 dark-subtracted 512x512 frames, decode-only prolix sweep, one stack frame-sharded over the GPUs",
                  w=512, h=512, dtype="int16", frames=8000, scaling="strong", decode_only=True, cpu_sample=3000, gen="dark"),
    "c5i32": dict(label="configs[4]: signed int32 dark-subtracted 4148x4362 frames, decode-only prolix sweep, one stack frame-sharded over the GPUs",
                  w=4148, h=4362, dtype="int32", frames=32, scaling="strong", decode_only=True, cpu_sample=8, gen="dark"),
}
LAMBDA, N_PEAKS = 2.0, 200


def metric_of(cfg):
    return ("%s frames/s, %dx%d %s stack (%s)" %
            ("prolix" if cfg["decode_only"] else "terse+prolix", cfg["w"], cfg["h"], cfg["dtype"],
             "decode every frame" if cfg["decode_only"] else "encode then decode every frame"))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--frames", type=int, default=0, help="c2: frames per GPU; other configs: frames of the whole stack (0: the config's own)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-chunk", type=int, default=1000,
                    help="frames per chunk of the streamed e2e (encode of chunk k+1 overlaps decode of chunk k)")
    ap.add_argument("--e2e-ramp", type=int, default=0, help="frames of the first and of the last chunk (0: --e2e-chunk)")
    ap.add_argument("--e2e-enc-threads", type=int, default=1)
    ap.add_argument("--e2e-dec-threads", type=int, default=1)
    ap.add_argument("--dropin-frames", type=int, default=2000, help="frames of the e2e through the drop-in class jpa::Terse (c2, N = 1)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="frames of the cpu_baseline sample (0: the config's own)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    a = ap.parse_args()
    a.cfg = CONFIGS[a.config]
    return a


# ------------------------------------------------------------------------------------------ inputs
def torch_dtype(torch, name):
    """torch carrier of a pixel type: the unsigned 16- and 32-bit patterns travel in int16 / int32 tensors."""
    return {"uint8": torch.uint8, "uint16": torch.int16, "int16": torch.int16, "uint32": torch.int32, "int32": torch.int32}[name]


def synth_stack(torch, cfg, f0, frames, dev, seed=1000):
    """Frames [f0, f0 + frames) of the config's stack as a (frames, h*w) tensor of the carrier type.  Every chunk of
    frames is seeded by its GLOBAL index, so a frame-sharded run sees the same stack whatever the number of GPUs."""
    import numpy as np
    H, W = cfg["h"], cfg["w"]
    N = H * W
    tdt = torch_dtype(torch, cfg["dtype"])
    out = torch.empty((frames, N), dtype=tdt, device=dev)
    chunk = max(1, min(200, (64 << 20) // N))
    g = torch.Generator(device=dev)
    gen = cfg["gen"]
    lo_frame = (f0 // chunk) * chunk
    for c0 in range(lo_frame, f0 + frames, chunk):
        g.manual_seed(seed + 7919 * (c0 // chunk))
        n = chunk
        if gen in ("bragg", "eiger"):
            lam, peaks, amp_hi, top = (LAMBDA, N_PEAKS, 3000.0, 65535.0) if gen == "bragg" else (0.5, 2000, 1.0e6, 4294967295.0)
            img = torch.poisson(torch.full((n, N), lam, device=dev), generator=g)
            r = torch.arange(-4, 5, device=dev, dtype=torch.float32)
            dy, dx = torch.meshgrid(r, r, indexing="ij")
            dy, dx = dy.reshape(1, 1, -1), dx.reshape(1, 1, -1)
            cy = torch.rand((n, peaks, 1), device=dev, generator=g) * (H - 1)
            cx = torch.rand((n, peaks, 1), device=dev, generator=g) * (W - 1)
            sigma = 1.0 + torch.rand((n, peaks, 1), device=dev, generator=g)
            amp = torch.exp(math.log(20.0) + torch.rand((n, peaks, 1), device=dev, generator=g) * math.log(amp_hi / 20.0))
            iy, ix = cy.round() + dy, cx.round() + dx
            val = torch.poisson(amp * torch.exp(-((iy - cy) ** 2 + (ix - cx) ** 2) / (2 * sigma * sigma)), generator=g)
            ok = (iy >= 0) & (iy < H) & (ix >= 0) & (ix < W)
            idx = (iy.clamp(0, H - 1).long() * W + ix.clamp(0, W - 1).long()).reshape(n, -1)   # (integers: H*W exceeds float32's 2^24)
            img.scatter_add_(1, idx, (val * ok).reshape(n, -1))
            img = img.clamp_(0, top).to(torch.int64).to(tdt)              # int -> int wraps: the unsigned bit pattern
        elif gen == "sparse":
            img = torch.poisson(torch.full((n, N), 0.02, device=dev), generator=g).to(torch.int32).to(tdt)
        else:                                                          # dark-subtracted: Poisson(3) - 3 + round(N(0, 2^2))
            img = (torch.poisson(torch.full((n, N), 3.0, device=dev), generator=g) - 3.0 +
                   torch.round(2.0 * torch.randn((n, N), device=dev, generator=g))).to(torch.int32).to(tdt)
        a, b = max(c0, f0), min(c0 + chunk, f0 + frames)
        if b > a:
            out[a - f0:b - f0] = img[a - c0:b - c0]
        del img
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


def np_code(name):
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    return orc.code_of(np.dtype(name))


class CpuCodec:
    """The reference CPU codec (oracle/_ref: the reference's own headers, -O3 -DNDEBUG; one jpa::Terse per frame, frames
    statically partitioned over `threads` std::threads) on a fixed sample; falls back to the C port (oracle/liboracle.so)
    when _ref was not built.  The output buffer is allocated and touched ONCE, outside every timed call: a fresh
    786 MB buffer per call made the round-1 reference arm pay first-touch page faults in its timed decode."""

    def __init__(self, px_np, threads, decode_only=False):
        import numpy as np
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import orc                                # TEST-ONLY checker: used here as the timed CPU baseline
        self.np, self.orc = np, orc
        self.px = np.ascontiguousarray(px_np)
        self.threads = threads
        self.decode_only = decode_only
        self.code = orc.code_of(self.px.dtype)
        self.out = np.zeros_like(self.px)         # touched here
        self.ref = orc.ref()
        self.kind = "reference" if self.ref is not None else "port"
        self.payload_bytes = 0
        if self.ref is None:
            orc.build()
            self.payloads = None

    def step(self):
        """one pass over the sample -> (encode seconds, decode seconds)"""
        np, orc = self.np, self.orc
        F, N = self.px.shape
        if self.ref is not None:
            tot = ctypes.c_size_t(0)
            te = 0.0
            if not self.decode_only or not self.payload_bytes:
                te = self.ref.ref_bench_encode(self.px.ctypes.data, self.code, N, F, self.threads, ctypes.byref(tot))
                self.payload_bytes = int(tot.value)
            td = self.ref.ref_bench_decode(self.px.ctypes.data, self.code, N, F, self.threads, self.out.ctypes.data)
            return (0.0 if self.decode_only else te), td
        from concurrent.futures import ThreadPoolExecutor
        signed = self.code >= orc.I8
        if self.payloads is None or not self.decode_only:
            payloads = [None] * F

            def enc(f):
                payloads[f] = orc.encode_frame(self.px[f])[0]

            with ThreadPoolExecutor(self.threads) as ex:
                t0 = time.perf_counter()
                list(ex.map(enc, range(F)))
                te = time.perf_counter() - t0
            self.payloads = payloads
            self.payload_bytes = sum(q.size for q in payloads)
        else:
            te = 0.0

        def dec(f):
            self.out[f] = orc.decode_frame(self.payloads[f], N, signed, self.px.dtype)[0]

        with ThreadPoolExecutor(self.threads) as ex:
            t0 = time.perf_counter()
            list(ex.map(dec, range(F)))
            td = time.perf_counter() - t0
        return (0.0 if self.decode_only else te), td

    def check(self):
        # (the reference decodes 32-bit-wide blocks into a same-width integer as zeros, SURVEY App. C5: the synthetic
        # stacks stay below that, so a plain comparison holds)
        assert self.np.array_equal(self.out, self.px), "reference CPU round trip failed"

    def run(self, warmup, steps):
        """-> mean (encode s, decode s) over `steps` timed passes after `warmup` untimed ones (the SAME estimator for
        cpu_baseline and for the --impl reference arm)"""
        for _ in range(warmup):
            self.step()
        ts = [self.step() for _ in range(steps)]
        self.check()
        return sum(t[0] for t in ts) / len(ts), sum(t[1] for t in ts) / len(ts)


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def synth_stack_cpu(cfg, frames, seed):
    """CPU-side frames of the config's distribution for the reference arm (no GPU needed): `frames` DISTINCT frames
    (a tiled handful of frames would sit in the CPU's caches and flatter the reference)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    H, W = cfg["h"], cfg["w"]
    dt = np.dtype(cfg["dtype"])
    gen = cfg["gen"]
    rng = np.random.default_rng(seed)
    if gen == "sparse":
        return rng.poisson(0.02, size=(frames, H * W)).astype(dt)
    if gen == "dark":
        return (rng.poisson(3.0, size=(frames, H * W)) - 3 + np.rint(2.0 * rng.standard_normal((frames, H * W)))).astype(dt)
    lam, peaks, amp_hi = (LAMBDA, N_PEAKS, 3000.0) if gen == "bragg" else (0.5, 2000, 1.0e6)
    out = np.empty((frames, H * W), dt)
    from concurrent.futures import ThreadPoolExecutor

    def one(f):                                                     # the oracle's C generator (splitmix64 + inverse-CDF Poisson + peaks)
        out[f] = orc.synth_frame(orc.code_of(dt), W, H, lam, peaks, seed + f, 20.0, amp_hi)

    with ThreadPoolExecutor(host_cores()) as ex:
        list(ex.map(one, range(frames)))
    return out


def cpu_sample_frames(a):
    n = a.cpu_sample or a.cfg["cpu_sample"]
    return min(n, a.frames) if a.frames else n


def run_reference(a):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores, same config,
    metric and unit as our arm; each step is one pass over a bounded sample of the workload (the same sample size and
    estimator as our arm's cpu_baseline)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    cfg = a.cfg
    cores = host_cores()
    sample = cpu_sample_frames(a)
    px = synth_stack_cpu(cfg, sample, 4242)
    N = px.shape[1]
    so = px.dtype.itemsize
    codec = CpuCodec(px, cores, cfg["decode_only"])
    te, td = codec.run(max(a.warmup, 1), a.steps)
    v = sample / (te + td)
    desc = "%d distinct synthetic frames of the config's distribution per step, %s, %d threads, mean of %d timed steps" % (
        sample, "decode only" if cfg["decode_only"] else "encode then decode", cores, a.steps)
    line = {
        "impl": "reference", "metric": metric_of(cfg), "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": max(a.warmup, 1), "ms_per_step": 1e3 * (te + td), "higher_is_better": True, "scaling": cfg["scaling"],
        "vs_baseline": None, "dtype": {"uint8": "u8", "uint16": "u16", "uint32": "u32", "int16": "i16", "int32": "i32"}[cfg["dtype"]],
        "data": "synthetic",
        "config": {"workload": cfg["label"], "config": a.config, "frame": "%dx%d" % (cfg["w"], cfg["h"]), "pixel": cfg["dtype"], "block": 12,
                   "sample": desc, "same_config": True},
        "encode_frames_per_s": None if cfg["decode_only"] else sample / te, "decode_frames_per_s": sample / td,
        "uncompressed_GBps": v * N * so / 1e9, "compression_ratio": codec.payload_bytes / (sample * N * float(so)),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": codec.kind, "sample": desc, "cpu": cpu_model()},
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
    cfg = a.cfg
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
    npdt = np.dtype(cfg["dtype"])
    so = npdt.itemsize
    signed = npdt.kind == "i"
    tcode = trpx_b200.dtype_code(npdt)
    N_VALUES = cfg["w"] * cfg["h"]
    decode_only = cfg["decode_only"]
    weak = cfg["scaling"] == "weak"
    total_frames = a.frames or cfg["frames"]
    if weak:                                                        # every rank its own stack
        F, f0, job_frames = total_frames, 0, world * total_frames
        seed = 1000 + 100000 * rank
    else:                                                           # ONE stack, contiguous frame ranges per rank
        q, r_ = divmod(total_frames, world)
        f0 = rank * q + min(rank, r_)
        F = q + (1 if rank < r_ else 0)
        job_frames = total_frames
        seed = 1000
        if F == 0:
            raise SystemExit("bench.py: fewer frames than GPUs")

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
    px = synth_stack(torch, cfg, f0, F, dev, seed)
    cap = trpx_b200.max_compressed_bytes(N_VALUES, npdt, 12, F)
    payload = torch.empty(cap, dtype=torch.uint8, device=dev)
    ends = torch.zeros(F, dtype=torch.int64, device=dev)
    small = torch.zeros(4, dtype=torch.int32, device=dev)          # prolix_bits, enc status, dec status
    back = torch.empty_like(px)
    stream = torch.cuda.current_stream().cuda_stream
    raw_bytes = F * N_VALUES * so

    def encode():
        codec.encode_device(px.data_ptr(), npdt, N_VALUES, F, payload.data_ptr(), cap, ends.data_ptr(),
                            small.data_ptr(), small.data_ptr() + 4, stream)

    def decode(nbytes):
        codec.decode_device(payload.data_ptr(), nbytes, signed, N_VALUES, F, ends.data_ptr(), back.data_ptr(),
                            npdt, small.data_ptr() + 8, stream, lane=1)

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
        if not decode_only:
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
        if not decode_only:
            encode()
        ev[k][1].record()
        decode(cbytes)
        ev[k][2].record()
        if a.steps <= 64:
            # per-kernel device times come from events the library drops between its kernels; reading
            # them needs a drained stream, so this costs one host sync per step (outside the kernels)
            torch.cuda.synchronize()
            for name, ms in (([] if decode_only else codec.last_kernel_times(0)) + codec.last_kernel_times(1)):
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

    # ---- single-frame latency (config[0]'s shape of call: one frame in, one frame out; device-resident)
    latency = None
    if rank == 0 and not a.no_e2e:
        lat_e, lat_d = [], []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for k in range(23):
            e0.record()
            codec.encode_device(px.data_ptr(), npdt, N_VALUES, 1, payload.data_ptr(), cap, ends.data_ptr(),
                                small.data_ptr(), small.data_ptr() + 4, stream, lane=2)
            e1.record()
            torch.cuda.synchronize()
            if k >= 3:
                lat_e.append(e0.elapsed_time(e1))
        cb1 = int(ends[0])
        for k in range(23):
            e0.record()
            codec.decode_device(payload.data_ptr(), cb1, signed, N_VALUES, 1, ends.data_ptr(), back.data_ptr(),
                                npdt, small.data_ptr() + 8, stream, lane=2)
            e1.record()
            torch.cuda.synchronize()
            if k >= 3:
                lat_d.append(e0.elapsed_time(e1))
        assert torch.equal(back[0], px[0])
        lat_e.sort()
        lat_d.sort()
        latency = {"what": "ONE %dx%d %s frame per call, device-resident, CUDA events around the call (median of 20)" % (cfg["w"], cfg["h"], cfg["dtype"]),
                   "encode_us": 1e3 * lat_e[len(lat_e) // 2], "decode_us": 1e3 * lat_d[len(lat_d) // 2]}
        encode()                                                    # restore the full stack's payload and frame ends
        torch.cuda.synchronize()

    # ---- e2e: host buffers (pinned), through the host-pointer C ABI; H2D + D2H inside the timed region
    e2e = None
    if not a.no_e2e:
        e2e = run_e2e(a, cfg, torch, np, trpx_b200, codec, dist, dev, local, rank, world, barrier, px, ends, F, N_VALUES, npdt, tcode,
                      signed, decode_only, job_frames)

    # ---- CPU baseline on this box's cores (rank 0, N == 1), and byte identity with the reference
    cpu = None
    identity = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import orc                                                  # TEST-ONLY checker
        cores = host_cores()
        sample = min(F, cpu_sample_frames(a))
        px_np = px[:sample].cpu().numpy().view(npdt)
        cc = CpuCodec(px_np, cores, decode_only)
        te, td = cc.run(1, 2)
        cpu = {"value": sample / (te + td), "unit": UNIT, "cores": cores, "kind": cc.kind,
               "sample": "first %d of the %d frames, %s, mean of 2 timed passes after one warm-up pass" % (
                   sample, F, "decode only" if decode_only else "encode then decode"),
               "encode_frames_per_s": None if decode_only else sample / te, "decode_frames_per_s": sample / td, "cpu": cpu_model()}
        del cc
        # byte identity: per-frame size and FNV-1a-64 of OUR payload against the reference's, on every frame the
        # reference can digest within about a minute (all 10,000 for configs[1])
        ref = orc.ref()
        if ref is not None:
            n_id = F if F * N_VALUES * so <= (6 << 30) else max(sample, 1)
            h_pay = payload[:int(ends[n_id - 1])].cpu().numpy()
            h_ends = ends[:n_id].cpu().numpy().astype(np.uint64)
            ours = np.zeros(n_id, np.uint64)
            orc.orc().orc_fnv64_frames(h_pay.ctypes.data, h_ends.ctypes.data, n_id, ours.ctypes.data)
            sizes = np.zeros(n_id, np.uint64)
            theirs = np.zeros(n_id, np.uint64)
            t0 = time.perf_counter()
            for c0 in range(0, n_id, 1000):                          # (host copies of the pixels in chunks)
                c1 = min(n_id, c0 + 1000)
                chunk = np.ascontiguousarray(px[c0:c1].cpu().numpy().view(npdt))
                rc = ref.ref_frame_digests(chunk.ctypes.data, orc.code_of(npdt), N_VALUES, c1 - c0, cores,
                                           sizes[c0:].ctypes.data, theirs[c0:].ctypes.data)
                assert rc == 0, "reference digest failed"
            our_sizes = np.diff(np.concatenate([[0], h_ends.astype(np.int64)])).astype(np.uint64)
            same = bool(np.array_equal(our_sizes, sizes) and np.array_equal(ours, theirs))
            assert same, "payload differs from the reference's (frames %s)" % np.nonzero((our_sizes != sizes) | (ours != theirs))[0][:8]
            identity = {"frames_compared": int(n_id), "of": int(F), "byte_identical": same,
                        "how": "per-frame payload size and FNV-1a-64 of every compared frame, ours vs the reference's (one jpa::Terse per frame)",
                        "seconds": time.perf_counter() - t0}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0
    peak, peak_src = peaks()
    alg_bytes = raw_bytes + cbytes                                  # per pass: N*sizeof(T) + C
    kavg = {n: sum(v) / len(v) for n, v in ktimes.items()}
    dom = max(kavg, key=kavg.get) if kavg else None
    traffic = {}
    tsrc = None
    for tf in ("r02_traffic.json", "r01_traffic.json"):             # DRAM bytes per launch from the committed ncu capture
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", tf)))
            if tj.get("frames") == F and tj.get("config", "c2") == a.config:
                traffic = {k: v["traffic_bytes"] for k, v in tj["kernels"].items()}
                tsrc = "profiles/%s (ncu --set full, same workload)" % tf
                break
        except Exception:
            pass
    enc_avg, dec_avg = enc_ms_max / a.steps, dec_ms_max / a.steps

    def pass_roof(ms, kernels):
        ach = alg_bytes / (ms * 1e-3) / 1e9
        tr = [traffic[k] for k in kernels if k in traffic]
        return {"ms": ms, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "frac_of_8TBps_nominal": ach / 8000.0,
                "traffic": sum(tr) if tr else None, "kernels": {k: kavg[k] for k in kernels if k in kavg}}

    enc_k = [k for k in kavg if k.startswith("terse")]
    dec_k = [k for k in kavg if k.startswith("prolix")]
    roof_passes = {"decode": pass_roof(dec_avg, dec_k)}
    if not decode_only:
        roof_passes["encode"] = pass_roof(enc_avg, enc_k)
    worst = min(roof_passes, key=lambda k: roof_passes[k]["frac"])
    roof = None
    if dom:
        ach = alg_bytes / (kavg[dom] * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic.get(dom), "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                "avg_launch_ms": kavg[dom], "frac_of_8TBps_nominal": ach / 8000.0, "traffic_source": tsrc if dom in traffic else None,
                # both passes (a pass = every kernel of one TERSE or PROLIX call): the worse one is the number to improve
                "passes": roof_passes, "worst_pass": worst, "worst_pass_frac": roof_passes[worst]["frac"]}
    value = job_frames * a.steps / (total_ms_max * 1e-3)
    sh = {"uint8": "u8", "uint16": "u16", "uint32": "u32", "int16": "i16", "int32": "i32"}[cfg["dtype"]]
    line = {
        "metric": metric_of(cfg), "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": total_ms_max / a.steps, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None,
        "dtype": sh, "data": "synthetic",
        "config": {"workload": cfg["label"], "config": a.config, "frames_per_gpu": F, "frames_whole_job": job_frames,
                   "frame": "%dx%d" % (cfg["w"], cfg["h"]), "pixel": cfg["dtype"], "block": 12, "generator": cfg["gen"],
                   "sharding": "every rank its own stack" if weak else "one stack, contiguous frame ranges per rank, no collective on the data path",
                   "l2": "inputs larger than L2 (%.2f GB of pixels per pass and rank vs 126 MB)" % (raw_bytes / 1e9)},
        "encode_frames_per_s": None if decode_only else job_frames / (enc_avg * 1e-3), "decode_frames_per_s": job_frames / (dec_avg * 1e-3),
        "encode_uncompressed_GBps": None if decode_only else job_frames * N_VALUES * so / (enc_avg * 1e-3) / 1e9,
        "decode_uncompressed_GBps": job_frames * N_VALUES * so / (dec_avg * 1e-3) / 1e9,
        "uncompressed_GBps": value * N_VALUES * so / 1e9,
        "compression_ratio": cbytes / float(raw_bytes), "prolix_bits": int(small[0]),
        "passes": {k: {"ms": v["ms"], "hbm_GBps": v["achieved"], "frac_of_measured_peak": v["frac"], "traffic": v["traffic"]} for k, v in roof_passes.items()},
        "kernel_ms": kavg, "roofline": roof, "cpu_baseline": cpu, "byte_identity": identity, "e2e": e2e, "latency": latency,
        "gpu_launches": int(launches), "clocks": clocks,
        "build": build_info(), "host_wall_s_timed_region": t_host1 - t_host0,
    }
    if a.config == "c2":
        line["vs_readme_claim_2000_frames_per_s"] = value / 2000.0
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def build_info():
    try:
        b = json.load(open(os.path.join(ROOT, "trpx_b200", "build_info.json")))
        return {"build_mode": b.get("build_mode"), "why": b.get("why"), "src_hash": b.get("src_hash", "")[:16]}
    except Exception:
        return None


E2E_RUNNER = None


def run_e2e(a, cfg, torch, np, trpx_b200, codec, dist, dev, local, rank, world, barrier, px, ends, F, N_VALUES, npdt, tcode, signed,
            decode_only, job_frames):
    """PCIe-inclusive: the same metric through the host-pointer C ABI with pinned host buffers; H2D and D2H inside the
    timed region.  configs[1] also runs the streamed / overlapped choreographies and, at N = 1, the drop-in class."""
    so = npdt.itemsize
    frame_raw = N_VALUES * so
    # the whole rank's stack in pinned host memory; should the host refuse that much (all ranks of a box pin at once),
    # the e2e runs on the first half, quarter ... of the frames and says so
    Fe = F
    while True:
        try:
            cb_e = int(ends[Fe - 1])
            pfs = (int(1.25 * cb_e / Fe) + 255) // 256 * 256   # payload slot bytes per frame of the streamed run
            h_px = torch.empty((Fe, N_VALUES), dtype=px.dtype, pin_memory=True)
            h_back = torch.empty((Fe, N_VALUES), dtype=px.dtype, pin_memory=True)
            h_payload = torch.empty(Fe * pfs + (Fe + 2) * 4096, dtype=torch.uint8, pin_memory=True)
            break
        except RuntimeError:
            h_px = h_back = h_payload = None
            if Fe <= 4:
                raise
            Fe //= 2
    if dist is not None:                                       # every rank runs the e2e on the same number of frames
        fe_t = torch.tensor([Fe], dtype=torch.int64, device=dev)
        dist.all_reduce(fe_t, op=dist.ReduceOp.MIN)
        if int(fe_t[0]) < Fe:
            Fe = int(fe_t[0])
            cb_e = int(ends[Fe - 1])
            h_px, h_back = h_px[:Fe], h_back[:Fe]
    raw_e = Fe * frame_raw
    h_px.copy_(px[:Fe])
    fb = np.zeros(Fe, np.uint64)
    L = trpx_b200.lib()
    tot = ctypes.c_size_t(0)
    pb = ctypes.c_uint(0)
    weak = cfg["scaling"] == "weak"
    e2e_job_frames = world * Fe if weak else (job_frames if Fe == F else world * Fe)
    seq_parts = []

    def sequential():
        """encode the whole stack, then decode the whole payload: two calls on one context"""
        t_a = time.perf_counter()
        if not decode_only or not seq_parts:
            rc = L.trpx_encode_host(codec._h, h_px.data_ptr(), tcode, N_VALUES, Fe, 12, h_payload.data_ptr(),
                                    h_payload.numel(), fb.ctypes.data, ctypes.byref(tot), ctypes.byref(pb))
            assert rc == 0, rc
        t_b = time.perf_counter()
        rc = L.trpx_decode_host(codec._h, h_payload.data_ptr(), tot.value, int(signed), 12, N_VALUES, Fe, 0, Fe,
                                fb.ctypes.data, None, h_back.data_ptr(), tcode)
        assert rc == 0, rc
        seq_parts.append((t_b - t_a, time.perf_counter() - t_b))
        return tot.value

    def timed(fn, decode_part_only=False):
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
                ts.append(seq_parts[-1][1] if decode_part_only else t1 - t0)
        t = torch.tensor(ts, dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ts = sorted(float(x) for x in t.cpu())
        return ts[len(ts) // 2] if len(ts) % 2 else 0.5 * (ts[len(ts) // 2 - 1] + ts[len(ts) // 2]), ts

    # ---- what this box can move: every rank copies the SAME bytes the sequential e2e moves, at the same time, with
    # plain large pinned copies and no codec at all.  (At N = 8 the ranks share one host's memory and PCIe roots: the
    # e2e number is then a fraction of this ceiling, not of 8 x the single-GPU link rate.)
    def host_ceiling():
        big = 128 << 20
        scratch = torch.empty(min(raw_e, 1 << 30), dtype=torch.uint8, device=dev)
        hp, hb = h_px.view(torch.uint8).reshape(-1), h_back.view(torch.uint8).reshape(-1)
        hpl = h_payload[:cb_e]

        def copy_all(dst_of, src_of, nbytes, h2d):
            for o in range(0, nbytes, big):
                n_ = min(big, nbytes - o)
                if h2d:
                    scratch[(o % scratch.numel()):(o % scratch.numel()) + n_].copy_(src_of[o:o + n_], non_blocking=True)
                else:
                    dst_of[o:o + n_].copy_(scratch[(o % scratch.numel()):(o % scratch.numel()) + n_], non_blocking=True)

        ts = []
        for k in range(3):
            barrier()
            t0 = time.perf_counter()
            if not decode_only:
                copy_all(None, hp, raw_e, True)                  # pixels up
                copy_all(hpl, None, cb_e, False)                 # payload down
            copy_all(None, hpl, cb_e, True)                      # payload up
            copy_all(hb, None, raw_e, False)                     # pixels down
            torch.cuda.synchronize()
            if k:
                ts.append(time.perf_counter() - t0)
        t = torch.tensor([min(ts)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        del scratch
        return float(t[0])

    ceil_s = host_ceiling()
    seq_s, seq_all = timed(sequential, decode_part_only=decode_only)
    h2d = (0 if decode_only else raw_e) + cb_e + 8 * Fe
    d2h = (0 if decode_only else cb_e) + raw_e + 16 * Fe
    e2e = {"value": e2e_job_frames / seq_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": 1e3 * seq_s, "uncompressed_GBps": e2e_job_frames * frame_raw / seq_s / 1e9, "frames_per_gpu": Fe,
           "api": ("trpx_decode_host" if decode_only else "trpx_encode_host + trpx_decode_host") + " (pinned host buffers)",
           "host_ceiling": {"value": e2e_job_frames / ceil_s, "ms_per_step": 1e3 * ceil_s, "e2e_over_ceiling": ceil_s / seq_s,
                            "how": "all ranks at once copy the bytes the sequential e2e moves (pixels up, payload down, payload up, pixels down) "
                                   "as plain 128 MB pinned copies, no codec; best of 2, max over ranks"},
           "mode": "sequential", "timing": "median of %d timed passes after one warm-up pass, host clock around the calls" % a.e2e_steps,
           "sequential": {"value": e2e_job_frames / seq_s, "ms_per_step": 1e3 * seq_s, "ms_all_steps": [round(1e3 * x, 2) for x in seq_all],
                          "encode_ms": None if decode_only else 1e3 * min(p_[0] for p_ in seq_parts[1:]),
                          "decode_ms": 1e3 * min(p_[1] for p_ in seq_parts[1:]),
                          "how": "one call encodes the whole stack, a second one decodes the whole payload (what a user of jpa::Terse does)"}}
    if a.config != "c2":
        return e2e

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

    def streamed():
        import queue
        qs = [queue.Queue() for _ in range(n_dec)]
        err = []
        sizes = [0] * n_chunks

        def enc_side(e):
            t_, p_ = ctypes.c_size_t(0), ctypes.c_uint(0)
            for c in range(e, n_chunks, n_enc):
                f0, nf = cuts[c], cuts[c + 1] - cuts[c]
                rc = L.trpx_encode_host(enc_ctx[e]._h, h_px.data_ptr() + f0 * frame_raw, tcode, N_VALUES, nf,
                                        12, h_slots.data_ptr() + slot_off[c], nf * pfs + 4096, fb.ctypes.data + 8 * f0,
                                        ctypes.byref(t_), ctypes.byref(p_))
                if rc != 0:
                    err.append(("encode", c, rc))
                sizes[c] = t_.value
                qs[c % n_dec].put((c, f0, nf, t_.value if rc == 0 else 0))

        def dec_side(d):
            for _ in range(d, n_chunks, n_dec):
                c, f0, nf, nbytes = qs[d].get()
                if not nbytes:
                    continue
                rc = L.trpx_decode_host(dec_ctx[d]._h, h_slots.data_ptr() + slot_off[c], nbytes, 0, 12, N_VALUES, nf, 0,
                                        nf, fb.ctypes.data + 8 * f0, None, h_back.data_ptr() + f0 * frame_raw, tcode)
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
            rc = L.trpx_encode_host(codec._h, h_px.data_ptr(), tcode, N_VALUES, Fe, 12, h_payload.data_ptr(),
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
                                            h_back.data_ptr() + f_prev * frame_raw, tcode)
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

    str_s, str_all = timed(streamed)
    ovl_s, ovl_all = timed(overlapped)
    for c_ in enc_ctx[1:] + dec_ctx:
        c_.close()
    # The headline e2e is the SEQUENTIAL pair of calls -- what a user of the drop-in class gets; the two pipelined
    # choreographies (several host threads and contexts) are reported beside it, not instead of it.
    e2e["pipelined_best"] = {"mode": "overlapped" if ovl_s <= str_s else "streamed", "value": e2e_job_frames / min(str_s, ovl_s),
                             "ms_per_step": 1e3 * min(str_s, ovl_s)}
    e2e["overlapped"] = {"value": e2e_job_frames / ovl_s, "ms_per_step": 1e3 * ovl_s,
                         "ms_all_steps": [round(1e3 * x, 2) for x in ovl_all], "min_decode_frames": chunk,
                         "how": "one trpx_encode_host call for the whole stack; a second host thread follows "
                                "trpx_ctx_encode_progress() and decodes the finished prefix on its own context"}
    e2e["streamed"] = {"value": e2e_job_frames / str_s, "ms_per_step": 1e3 * str_s, "ms_all_steps": [round(1e3 * x, 2) for x in str_all],
                       "chunk_frames": chunk, "first_last_chunk_frames": ramp, "encoder_threads": n_enc, "decoder_threads": n_dec,
                       "how": "host threads with one context each: chunks are encoded round-robin and decoded as "
                              "soon as they are encoded, so H2D and D2H overlap (full-duplex PCIe)"}
    # ---- through the drop-in class itself (pageable std::vectors, the reference's call sequence): N = 1 only
    if world == 1 and rank == 0:
        e2e["dropin"] = dropin_e2e(a, np, px, N_VALUES, min(F, a.dropin_frames))
    del h_px, h_back, h_payload
    return e2e


def dropin_e2e(a, np, px, n_values, frames):
    """jpa::Terse::push_back_frames + prolix_frames on pageable std::vectors (cxx/terse_bench.cpp): the class as it
    ships (payload through a pinned scratch buffer, caller ranges left pageable) and with per-call page-locking of the
    caller's ranges (TRPX_PIN_MIN_MB), which costs more than it saves."""
    exe = os.path.join(ROOT, "cxx", "terse_bench")
    if not os.path.exists(exe):
        return {"unavailable": "cxx/terse_bench has not been built"}
    raw = "/dev/shm/trpx_bench_%d.raw" % os.getpid()
    try:
        px[:frames].cpu().numpy().tofile(raw)
        out = {}
        for key, env in (("default", {}), ("caller_ranges_pinned_per_call", {"TRPX_PIN_MIN_MB": "64"})):
            e = dict(os.environ)
            e.update(env)
            r = subprocess.run([exe, raw, str(n_values), str(frames), "3"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=e, timeout=600)
            if r.returncode != 0:
                out[key] = {"failed": (r.stderr or r.stdout)[-300:]}
                continue
            out[key] = json.loads(r.stdout.strip().splitlines()[-1])
        return out
    finally:
        try:
            os.remove(raw)
        except OSError:
            pass


def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)
    return run_ours(a)


if __name__ == "__main__":
    sys.exit(main())
