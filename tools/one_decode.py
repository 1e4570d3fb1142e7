"""A few device-resident decodes of BASELINE configs[1] (48 kHz x 900 s), for ncu captures (tools/ncu_capture.sh)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import noaa_apt_b200 as na
from noaa_apt_b200 import synth

rate = int(sys.argv[1]) if len(sys.argv) > 1 else 48000
seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 900.0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
x = torch.from_numpy(synth.apt_signal(rate, seconds, seed=0)).cuda()
with na.Decoder(rate, na.Settings(), max_samples=x.numel()) as dec:
    bound = dec.out_bound(x.numel())
    out = torch.empty(bound, dtype=torch.float32, device="cuda")
    for _ in range(reps):
        dec.submit_device(x.data_ptr(), na._lib.F32, x.numel(), True, out.data_ptr(), bound)
        n = dec.wait()
    print("rows", n // 2080, dec.last_counts())
