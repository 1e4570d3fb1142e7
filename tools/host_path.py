"""apt_decode() on pageable buffers, a few calls, with APTB200_TRACE_HOST=1 timings (run on the GPU box)."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import noaa_apt_b200 as na

lib = na._lib.load()
lib.apt_bind_thread_to_device(0)
cache = "/tmp/apt_synth_48000_900.npy"
if os.path.exists(cache):
    pcm = np.load(cache)
else:
    from noaa_apt_b200 import synth
    pcm = synth.apt_pcm16(48000, 900, seed=0)
    np.save(cache, pcm)
x = pcm.astype(np.float32)
s = na.Settings().to_c()
bound = na.decode_len_bound(x.size, 48000)
out = np.zeros(bound, dtype=np.float32)
nout = C.c_uint64(0)
cb0 = na._lib.STATUS_CB()
for fmt, buf, name in ((na._lib.F32, x, "f32"), (na._lib.PCM16, pcm, "pcm16")):
    fn = lib.apt_decode if fmt == na._lib.F32 else lib.apt_decode_pcm16
    ts = []
    for i in range(6):
        t0 = time.perf_counter()
        rc = fn(buf.ctypes.data, buf.size, 48000, C.byref(s), 1, out.ctypes.data, bound, C.byref(nout), cb0, None)
        ts.append((time.perf_counter() - t0) * 1e3)
        assert rc == 0
    print(f"apt_decode {name}: " + " ".join(f"{t:.2f}" for t in ts) + " ms", flush=True)
