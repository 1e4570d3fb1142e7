#!/usr/bin/env python
"""bench.py -- throughput of the APT decode hot path on B200 (BASELINE.json metric).

One "step" = one pass of decode() (resample -> envelope -> low-pass -> sync -> rows) over one
synthetic recording per GPU.  Workload at every N: BASELINE.json configs[1], a single synthetic
48 kHz, 15-min, 2.4 kHz-subcarrier APT recording (43.2 M samples) per GPU ("weak" scaling: each
rank decodes its own recording, no collective on the data path).

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference        # CPU arm: the oracle port of the Rust reference

JSON keys beyond the base contract: roofline (dominant kernel, CUDA-event timed on the decoder's
stream), cpu_baseline (oracle on the host cores), e2e (host buffers through the C ABI), clocks.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "input_msamples_per_s_decoded"
UNIT = "Msamples/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rate", type=int, default=48000, help="input sample rate (Hz)")
    ap.add_argument("--seconds", type=float, default=900.0, help="recording length")
    ap.add_argument("--cpu-seconds", type=float, default=300.0,
                    help="length of the recording slice the CPU arms decode per step")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3"],
                    help="c2: BASELINE configs[1], one 48 kHz 15-min recording per GPU (default, the metric's config); "
                         "c3: configs[2], one 96 kHz 10-hour recording (a 900-s synthetic recording repeated 40x), "
                         "uploaded in overlapping chunks for the end-to-end number")
    ap.add_argument("--batch", type=int, default=1,
                    help="recordings decoded per GPU per step, one per CUDA stream (BASELINE configs[3]: 64)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                return float(json.load(f)["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """SM clock and clock-event (throttle) reasons sampled WHILE a timed region runs (B200_PROFILING.md): NVML through
    nvidia_ml_py polled every ~1 ms by a thread that only records while `active` is set (timed() sets it around each
    timed region); `nvidia-smi -lms` as the fallback when NVML cannot be loaded."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.index = index
        self.sm, self.reasons, self.mx = [], set(), None
        self.active = False
        self.running = False
        self.thread = None
        self.proc = None
        self.rows = []
        self.source = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.source = "nvml"
            self.running = True
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while self.running:
            if self.active:
                try:
                    self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                    mask = int(get_reasons(self.handle))
                    for name, bit in self.REASONS:
                        if mask & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(0.001)

    def _read(self):
        for line in self.proc.stdout:
            if self.active:
                self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.source == "nvml":
            self.running = False
            self.thread.join(timeout=1)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml, polled inside the timed regions"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 20 inside the timed regions"}


def make_recording(rate, seconds, seed):
    from noaa_apt_b200 import synth
    return synth.apt_pcm16(rate, seconds, seed=seed)


def cpu_arm(pcm, rate, steps, warmup):
    """Times the CPU oracle (C restatement of the Rust reference, 1 thread like the reference)."""
    import oracle
    x = pcm.astype(np.float32)
    for _ in range(warmup):
        oracle.decode(x, rate)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle.decode(x, rate)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return x.size / dt / 1e6, dt


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port: the Rust crate cannot be built
    here) on the host cores, same config/metric, each step a bounded slice of the workload."""
    if rank != 0:
        return
    pcm = make_recording(args.rate, min(args.cpu_seconds, args.seconds), seed=0)
    value, dt = cpu_arm(pcm, args.rate, args.steps, min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"single synthetic {args.rate} Hz {args.seconds:g}-s APT recording (BASELINE configs[1])",
                   "profile": "standard"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"first {min(args.cpu_seconds, args.seconds):g} s of the recording per step; "
                                   f"C restatement of the single-threaded Rust decode (no Rust toolchain in the image)",
                         "host_cores_available": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_b200(args, rank, local_rank, world):
    import torch
    import noaa_apt_b200 as na

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist_on = world > 1
    if dist_on:
        os.environ.setdefault("NCCL_DEBUG", "WARN")   # keep NCCL's version banner off stdout: one JSON line only
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from noaa_apt_b200 import sharding
    repeat = 1
    if args.workload == "c3":
        args.rate, args.seconds, args.batch, repeat = 96000, 900.0, 1, 40
    rate, K, W, B = args.rate, args.steps, max(args.warmup, 3), max(args.batch, 1)
    dev = f"cuda:{local_rank}"
    settings = na.Settings()
    # B recordings per GPU, one decoder (= one CUDA stream + workspaces) each; up to 4 distinct seeds are
    # generated and replicated into separate device buffers
    n_seeds = min(B, 4)
    pcms = [make_recording(rate, args.seconds, seed=rank * 4 + k) for k in range(n_seeds)]
    if repeat > 1:
        # 900 s = 1800 whole lines and 2 160 000 carrier cycles: the repetition is a continuous APT signal
        pcms = [np.tile(p, repeat) for p in pcms]
    n = pcms[0].size
    decs = [na.Decoder(rate, settings, max_samples=n, device=local_rank) for _ in range(B)]
    dec = decs[0]
    bound = dec.out_bound(n)

    # ---- device-resident arm ("value"): the Signals (f32, wav.rs:37) already in HBM ----
    x_hosts = [torch.from_numpy(p.astype(np.float32)).pin_memory() for p in pcms]
    x_devs = [x_hosts[k % n_seeds].to(dev) if k < n_seeds else x_hosts[k % n_seeds].to(dev).clone() for k in range(B)]
    out_devs = [torch.empty(bound, dtype=torch.float32, device=dev) for _ in range(B)]
    stream = torch.cuda.ExternalStream(dec.stream, device=local_rank)

    def step_device():
        for k in range(B):
            decs[k].submit_device(x_devs[k].data_ptr(), na._lib.F32, n, True, out_devs[k].data_ptr(), bound)
        got = 0
        for k in range(B):
            got = decs[k].wait()
        return got

    def barrier():
        torch.cuda.synchronize()
        if dist_on:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize; device time between two events recorded on decoder 0's
        stream after/before full-device synchronisation, so it spans the work of every stream."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = sum(d.launch_count for d in decs)
        sampler.active = True
        e0.record(stream)
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        e1.record(stream)
        e1.synchronize()
        sampler.active = False
        ms = e0.elapsed_time(e1)
        return ms, sum(d.launch_count for d in decs) - l0

    sampler = ClockSampler(local_rank)      # records only inside timed(): the device-resident and the two end-to-end legs
    sampler.start()
    for _ in range(W):
        produced = step_device()
    ms_local, launches = timed(step_device, K)
    value, ms_total = sharding.aggregate_throughput(B * n * K, ms_local, dist if dist_on else None, dev)
    ms_step = ms_total / K

    # ---- end-to-end arm: host buffers through the reference-facing call (H2D + D2H inside) ----
    out_hosts = [torch.empty(bound, dtype=torch.float32).pin_memory() for _ in range(B)]

    def step_host():
        for k in range(B):
            decs[k].submit_host_ptr(x_hosts[k % n_seeds].data_ptr(), na._lib.F32, n, True, out_hosts[k].data_ptr(), bound)
        got = 0
        for k in range(B):
            got = decs[k].wait()
        return got

    for _ in range(2):
        produced_host = step_host()
    e2e_local, _ = timed(step_host, K)
    e2e_value, e2e_total = sharding.aggregate_throughput(B * n * K, e2e_local, dist if dist_on else None, dev)
    e2e_ms = e2e_total / K

    # same, handing over the WAV's PCM16 samples (apt_decode_pcm16's path: the `as f32` of wav.rs:37 runs on the GPU)
    p_hosts = [torch.from_numpy(p.copy()).pin_memory() for p in pcms]

    def step_host_pcm16():
        for k in range(B):
            decs[k].submit_host_ptr(p_hosts[k % n_seeds].data_ptr(), na._lib.PCM16, n, True, out_hosts[k].data_ptr(), bound)
        for k in range(B):
            decs[k].wait()

    for _ in range(2):
        step_host_pcm16()
    p16_local, _ = timed(step_host_pcm16, K)
    p16_value, p16_total = sharding.aggregate_throughput(B * n * K, p16_local, dist if dist_on else None, dev)
    clocks = sampler.stop()

    # ---- roofline of the dominant kernel: CUDA events on the decoder's stream, per launch ----
    dec.set_profiling(True)
    acc = {}
    prof_steps = min(K, 10)
    for _ in range(prof_steps):
        # one recording alone on the GPU: the per-kernel times are not disturbed by the other streams
        dec.submit_device(x_devs[0].data_ptr(), na._lib.F32, n, True, out_devs[0].data_ptr(), bound)
        dec.wait()
        for name, ms in dec.kernel_times_ms():
            acc.setdefault(name, []).append(ms)
    dec.set_profiling(False)
    kernel_ms = {k: float(np.mean(v)) for k, v in acc.items()}
    counts = dec.last_counts()
    n_work = counts["n_work"]
    peak, peak_kind = measured_peaks()
    dom = "resample_envelope"
    alg_bytes = 4 * n + 4 * n_work                      # read every input once, write every e once
    achieved = alg_bytes / (kernel_ms[dom] * 1e-3) / 1e9 if dom in kernel_ms else None
    traffic = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of that kernel, from the committed ncu --set full capture
        with open(os.path.join(ROOT, "profiles", "r01_ncu_all_kernels_metrics.json")) as f:
            for m in json.load(f):
                if "k_polyphase_ut" in m["kernel"] and abs(args.seconds - 900.0) < 1e-6 and rate == 48000:
                    traffic = (m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]) * 1e6
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak if achieved else None, "traffic": traffic, "peak_source": peak_kind,
                "algorithmic_bytes": alg_bytes, "kernel_ms": kernel_ms.get(dom), "kernel_name": "k_polyphase_ut" if 12480 // math.gcd(rate, 12480) == 13 else "k_polyphase_ws / k_polyphase_generic",
                "all_kernels_ms": kernel_ms}

    line = None
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            sl = pcms[0][: int(min(args.cpu_seconds, args.seconds * repeat) * rate)]
            v, dt = cpu_arm(sl, rate, steps=3, warmup=1)
            cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"first {sl.size / rate:g} s of the same recording, 3 timed decodes; C restatement of the "
                             f"single-threaded Rust decode", "host_cores_available": os.cpu_count()}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": (f"single synthetic {rate} Hz {args.seconds * repeat:g}-s APT recording "
                                    f"(BASELINE configs[2]; 900-s recording x{repeat}; host upload chunked with "
                                    f"filter-length overlap)" if args.workload == "c3" else
                                    f"single synthetic {rate} Hz {args.seconds:g}-s APT recording per GPU "
                                    f"(BASELINE configs[1])" if B == 1 else
                                    f"batch of {B} synthetic {rate} Hz {args.seconds:g}-s APT recordings per GPU, one per "
                                    f"CUDA stream (BASELINE configs[3])"),
                       "profile": "standard", "recordings_per_gpu": B, "samples_per_recording": int(n),
                       "work_samples": int(n_work), "rows": int(produced // 2080), "sync_roots": int(counts["n_roots"]),
                       "l2": f"inputs_exceed_l2 ({4 * n / 1e6:.1f} MB f32 input per recording > 126 MB L2)",
                       "sharding": "one recording per GPU, no collective"},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(4 * n * B),
                    "d2h_bytes_per_step": int((4 * produced_host + 32) * B), "input": "pinned host f32 Signal"},
            "e2e_pcm16": {"value": p16_value, "unit": UNIT, "ms_per_step": p16_total / K, "h2d_bytes_per_step": int(2 * n * B),
                          "input": "pinned host int16 (the WAV's samples); cast on the device"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    for d in decs:
        d.close()
    if dist_on:
        dist.barrier()
        dist.destroy_process_group()
    return line


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
