"""Config 1 of BASELINE.json: decode the reference's own test recording on the
CPU (plumbing run).  The reference cannot be executed (no Rust toolchain), so
this checks the oracle against the independent numpy-f32 cross-check numbers
recorded in SURVEY.md Appendix C.  Runs only where /root/reference is mounted
(the build container); on the GPU box the file does not exist and the test is
skipped -- nothing in the `-m gpu` suite reads /root/reference.
"""
import os
import wave

import numpy as np
import pytest

import oracle

WAV = "/root/reference/test/test_11025hz.wav"


@pytest.mark.skipif(not os.path.exists(WAV), reason="reference test recording not mounted")
def test_decode_reference_recording_matches_survey_numbers():
    with wave.open(WAV) as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate()) == (1, 2, 11025)
        n = w.getnframes()
        pcm = np.frombuffer(w.readframes(n), dtype="<i2")
    assert n == 9_067_017
    x = oracle.pcm16_to_f32(pcm)
    out, st = oracle.decode_steps(x, 11025)
    assert st["resampled"].size == 10_263_607
    assert st["sync_pos"].size == 1644
    assert st["sync_pos"][:4].tolist() == [3989, 13153, 19644, 26047]
    d = np.diff(st["sync_pos"].astype(np.int64))
    assert (np.median(d), d.min(), d.max()) == (6240, 0, 18708)
    assert out.size == 1642 * 2080
    assert abs(float(st["resampled"].min()) - (-26.86)) < 0.01
    assert abs(float(st["resampled"].max()) - 32.86) < 0.01
    assert abs(float(st["demodulated"].max()) - 67.60) < 0.01
    assert abs(float(out.min()) - (-3.02)) < 0.01 and abs(float(out.max()) - 44.78) < 0.01
    assert abs(float(out.mean(dtype=np.float64)) - 9.38) < 0.01
