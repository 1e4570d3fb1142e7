"""Generates tests/golden/decode_golden.npz with the CPU oracle (oracle/apt_oracle.c).

The reference itself cannot be run (Rust, no toolchain), so these vectors pin the CUDA path and
the oracle against drift, not against a run of the reference.  Re-run only deliberately:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from noaa_apt_b200 import synth  # noqa: E402

out = {}
for rate, seconds, seed in ((11025, 8.0, 101), (48000, 8.0, 102), (96000, 8.0, 103)):
    pcm = synth.apt_pcm16(rate, seconds, seed=seed)
    rows, steps = oracle.decode_steps(pcm.astype(np.float32), rate)
    out[f"rows_{rate}"] = rows
    out[f"sync_{rate}"] = steps["sync_pos"]
    out[f"seconds_{rate}"] = np.float64(seconds)
    out[f"seed_{rate}"] = np.int64(seed)
    # a cheap fingerprint of the input so a change of the generator is noticed
    out[f"pcm_sum_{rate}"] = np.int64(pcm.astype(np.int64).sum())
    print(rate, rows.size // 2080, "rows", steps["sync_pos"][:4])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "decode_golden.npz"), **out)
