"""Per-kernel device times (CUDA events on the decoder's stream) of one decode of a synthetic recording.
    python tools/kernel_times.py [rate] [seconds] [reps]         (environment switches select kernel variants)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import noaa_apt_b200 as na
from noaa_apt_b200 import synth

rate = int(sys.argv[1]) if len(sys.argv) > 1 else 48000
seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 900.0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
cache = f"/tmp/apt_synth_{rate}_{int(seconds)}.npy"
if os.path.exists(cache):
    pcm = np.load(cache)
else:
    pcm = synth.apt_pcm16(rate, seconds, seed=0)
    np.save(cache, pcm)
x = torch.from_numpy(pcm.astype(np.float32)).cuda()
with na.Decoder(rate, na.Settings(), max_samples=x.numel()) as dec:
    bound = dec.out_bound(x.numel())
    out = torch.empty(bound, dtype=torch.float32, device="cuda")
    for _ in range(3):
        dec.submit_device(x.data_ptr(), na._lib.F32, x.numel(), True, out.data_ptr(), bound)
        dec.wait()
    dec.set_profiling(True)
    acc = {}
    for _ in range(reps):
        dec.submit_device(x.data_ptr(), na._lib.F32, x.numel(), True, out.data_ptr(), bound)
        n = dec.wait()
        for name, ms in dec.kernel_times_ms():
            acc.setdefault(name, []).append(ms * 1e3)
    dec.set_profiling(False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s = torch.cuda.ExternalStream(dec.stream)
    e0.record(s)
    for _ in range(reps):
        dec.submit_device(x.data_ptr(), na._lib.F32, x.numel(), True, out.data_ptr(), bound)
        dec.wait()
    e1.record(s)
    e1.synchronize()
tags = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("APTB200_"))
print(f"[{tags or 'default'}] {rate} Hz {seconds:g} s rows {n // 2080}: " +
      "  ".join(f"{k} {np.median(v):.1f}" for k, v in acc.items()) +
      f"  | sum {sum(np.median(v) for v in acc.values()):.1f} us, decode {e0.elapsed_time(e1) / reps * 1e3:.1f} us")
