"""Small decodes through every resampler path, for compute-sanitizer (tools/sanitize.sh)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import noaa_apt_b200 as na
from noaa_apt_b200 import synth
import oracle

cases = [(48000, "standard", 12), (96000, "standard", 12), (11025, "standard", 14), (48000, "fast", 12), (24960, "standard", 12)]
if len(sys.argv) > 1:
    cases = cases[: int(sys.argv[1])]
for rate, profile, seconds in cases:
    x = synth.apt_signal(rate, seconds, seed=3)
    s = na.Settings.profile(profile)
    with na.Decoder(rate, s, max_samples=x.size) as dec:
        got = dec.decode(x)
        pos = dec.last_sync()
        got16 = dec.decode(synth.apt_pcm16(rate, seconds, seed=3))
    os_ = oracle.default_settings()
    os_.work_rate, os_.resample_atten = s.work_rate, s.resample_atten
    os_.resample_delta_freq, os_.resample_cutout, os_.demodulation_atten = s.resample_delta_freq, s.resample_cutout, s.demodulation_atten
    ref, st = oracle.decode_steps(x, rate, os_)
    ok = np.array_equal(pos, st["sync_pos"]) and got.size == ref.size and got16.size == ref.size
    print(rate, profile, "rows", got.size // 2080, "sync equal" if ok else "MISMATCH", flush=True)
