"""apt_decode_batch over every GPU of the box (BASELINE configs[4] in miniature): recording i -> device i % G, one feeder
thread per device, no inter-device communication; every recording against the oracle.  Skipped on a one-GPU box."""
import numpy as np
import pytest

import noaa_apt_b200 as na
from noaa_apt_b200 import synth
import oracle

pytestmark = pytest.mark.gpu


def nerr(got, ref):
    assert got.shape == ref.shape, (got.shape, ref.shape)
    scale = float(np.max(np.abs(ref))) if ref.size else 1.0
    return float(np.max(np.abs(got.astype(np.float64) - ref.astype(np.float64)))) / (scale or 1.0)


def test_decode_batch_on_all_devices():
    g = na.device_count()
    if g < 2:
        pytest.skip("needs more than one GPU")
    rate = 48000
    base = [synth.apt_pcm16(rate, 40 + 3 * k, seed=40 + k) for k in range(4)]       # different lengths
    refs = [oracle.decode(p.astype(np.float32), rate) for p in base]
    count = 6 * g + 3                                                                  # uneven over the devices
    sigs = [base[i % 4].astype(np.float32) for i in range(count)]
    outs, statuses = na.decode_batch(sigs, rate, na.Settings(), True, devices=list(range(g)), streams_per_device=3)
    assert statuses == [0] * count
    for i, got in enumerate(outs):
        ref = refs[i % 4]
        assert got.size == ref.size and nerr(got, ref) <= 1e-5, i
    # PCM16 input, and a recording that is too short in the middle of the batch: its status, nobody else's
    sigs16 = [base[i % 4] for i in range(2 * g)]
    sigs16[g] = np.zeros(20000, np.int16)
    outs, statuses = na.decode_batch(sigs16, rate, na.Settings(), True, devices=list(range(g)), streams_per_device=2)
    assert statuses[g] == na._lib.ERR_TOO_SHORT and outs[g] is None
    for i, got in enumerate(outs):
        if i != g:
            assert statuses[i] == 0 and nerr(got, refs[i % 4]) <= 1e-5
    # a device that does not exist: every recording of that device reports the failure, the others decode
    outs, statuses = na.decode_batch(sigs[:2 * g], rate, na.Settings(), True, devices=list(range(g - 1)) + [99], streams_per_device=2)
    for i in range(2 * g):
        if i % g == g - 1:
            assert statuses[i] != 0 and outs[i] is None
        else:
            assert statuses[i] == 0
