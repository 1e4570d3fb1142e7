"""CPU-only checks of the drop-in boundary: libaptb200.so loads without a GPU, exports every symbol
include/aptb200.h declares, and its host-side logic (filter design, unit arithmetic, length and error
rules) agrees with the oracle.  No kernel runs here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import __graft_entry__ as entry

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def na():
    entry.build()
    import noaa_apt_b200
    return noaa_apt_b200


def declared_functions():
    text = open(os.path.join(ROOT, "include", "aptb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(apt_[a-z0-9_]+)\s*\(", text))
    names.discard("apt_status_cb")
    return sorted(names)


def test_library_exports_every_declared_symbol(na):
    lib = C.CDLL(na.library_path())
    names = declared_functions()
    assert len(names) >= 40
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in aptb200.h but not exported"
        assert name in na._lib.SIGNATURES, f"{name} has no ctypes prototype"
    assert set(na._lib.SIGNATURES) == set(names)
    assert lib.apt_abi_version() == 1


def test_filter_taps_bit_identical_to_oracle(na):
    import oracle
    for rate, l in ((11025, 832), (48000, 13), (96000, 13), (44100, 208)):
        f = na.filters.LowpassDcRemoval(na.Freq.hz(4800, rate), 30.0, na.Freq.hz(1000, rate))
        f.resample(rate, rate * l)
        ratio = np.float32(rate * l) / np.float32(rate)
        ref = oracle.design(oracle.FILTER_LOWPASS_DC, np.float32(oracle.freq_hz(4800, rate)) / ratio, 30.0,
                            np.float32(oracle.freq_hz(1000, rate)) / ratio)
        got = f.design()
        assert got.size == ref.size
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    assert [14057, 959, 1915][1] == 959
    cut = np.float32(4160) / np.float32(12480)
    lp = na.filters.Lowpass(na.Freq.pi_rad(cut), 25.0, na.Freq.pi_rad(cut) / 5.0).design()
    assert lp.size == 37
    assert np.array_equal(lp, oracle.design(oracle.FILTER_LOWPASS, cut, 25.0, cut / np.float32(5)))
    assert na.filters.NoFilter().design().tolist() == [1.0]


def test_tap_counts_match_survey_appendix_b(na):
    # 14057 / 959 / 1915 taps for the standard profile
    for rate, l, n in ((11025, 832, 14057), (48000, 13, 959), (96000, 13, 1915)):
        f = na.filters.LowpassDcRemoval(na.Freq.hz(4800, rate), 30.0, na.Freq.hz(1000, rate))
        f.resample(rate, rate * l)
        assert f.design().size == n


def test_filter_resample_equivalence(na):
    # filters.rs:377-413
    f = na.filters.Lowpass(na.Freq.hz(123.0, 1000), 40.0, na.Freq.hz(12.0, 1000))
    f.resample(na.Rate.hz(1000), na.Rate.hz(3000))
    assert f == na.filters.Lowpass(na.Freq.hz(123.0, 3000), 40.0, na.Freq.hz(12.0, 3000))
    g = na.filters.NoFilter()
    g.resample(1000, 3000)
    assert g == na.filters.NoFilter()


def test_bessel_and_sync_frame(na):
    import oracle
    lib = na._lib.load()
    for x in np.linspace(0, 7, 15):
        assert lib.apt_bessel_i0(float(x)) == oracle.bessel_i0(float(x))
    for wr in (4160 * 2, 4160 * 3, 4160 * 5):
        assert np.array_equal(na.generate_sync_frame(wr), oracle.generate_sync_frame(wr))
    with pytest.raises(na.err.Internal) as e:
        na.generate_sync_frame(11025)
    assert e.value.code == na._lib.ERR_WORK_RATE


def test_lengths_and_errors_without_a_gpu(na):
    import oracle
    lib = na._lib.load()
    # exact output length of resample_with_filter (polyphase and L == 1)
    rng = np.random.default_rng(0)
    for in_rate, out_rate, n in ((11025, 12480, 5000), (48000, 12480, 9000), (24960, 12480, 7001), (1000, 1500, 100)):
        cf = na.filters.Lowpass(na.Freq.pi_rad(0.2), 30.0, na.Freq.pi_rad(0.05)).to_c()
        got = C.c_uint64(0)
        assert lib.apt_resample_len(n, in_rate, out_rate, C.byref(cf), C.byref(got)) == 0
        ref = oracle.resample_with_filter(rng.standard_normal(n).astype(np.float32), in_rate, out_rate,
                                          oracle.FILTER_LOWPASS, 0.2, 30.0, 0.05)
        assert got.value == ref.size
    # dsp.rs:420-434: RateOverflow for two prime rates
    cf = na.filters.NoFilter().to_c()
    got = C.c_uint64(0)
    assert lib.apt_resample_len(1000, 99371, 93911, C.byref(cf), C.byref(got)) == na._lib.ERR_RATE_OVERFLOW
    assert b"divisor" in lib.apt_last_error()
    # dsp.rs:69-71
    assert lib.apt_resample_len(1000, 11025, 0, C.byref(cf), C.byref(got)) == na._lib.ERR_RESAMPLE_TO_ZERO
    # decode.rs:79-83 is decided from lengths alone, before any device work
    with pytest.raises(na.err.Internal) as e:
        na.decode(na.Context(), na.Settings(), np.zeros(20000, np.float32), 11025, True)
    assert e.value.code == na._lib.ERR_TOO_SHORT
    with pytest.raises(na.err.RateOverflow):
        na.decode(na.Context(), na.Settings(work_rate=93911), np.zeros(200000, np.float32), 99371, True)
    with pytest.raises(na.err.InvalidInput):
        na.decode(na.Context(), na.Settings(), np.zeros(0, np.float32), 11025, True)
    # bound = rows * 2080
    assert na.decode_len_bound(9_067_017, 11025) == (10_263_607 // 6240) * 2080


def test_compute_fails_loudly_without_a_device(na):
    if na.device_count() > 0:
        pytest.skip("a CUDA device is present")
    from noaa_apt_b200 import synth
    x = synth.apt_signal(11025, 8, seed=1)
    with pytest.raises(na.err.CudaError) as e:
        na.decode(na.Context(), na.Settings(), x, 11025, True)
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(na.err.CudaError):
        na.dsp.demodulate(na.Context(), x[:100], na.Freq.hz(2400, 12480))
    with pytest.raises(na.err.CudaError):
        na.Decoder(11025, na.Settings(), max_samples=x.size)


def test_profiles(na):
    s = na.Settings.profile("fast")
    assert (s.work_rate, s.resample_delta_freq, s.demodulation_atten) == (16640, 3000.0, 23.0)
    s = na.Settings.profile("slow")
    assert (s.work_rate, s.resample_atten, s.resample_delta_freq) == (20800, 40.0, 500.0)
    assert na.Settings.profile("standard") == na.Settings()


def _tile_plan(na, l, m, taps):
    """(info, tapsA[g][u][4], tapsB[g][u][4], w0[g]) with u relative to w0 (A) -- B's u also relative to w0."""
    lib = na._lib.load()
    info = na._lib.CTileInfo()
    assert lib.apt_tile_plan(l, m, taps.ctypes.data, taps.size, C.byref(info), None, 0, None, 0) == 0
    if not info.usable:
        return None, None, None, None
    tt = np.zeros(info.groups * info.group_stride, np.float32)
    xs = np.zeros(info.groups, np.uint32)
    assert lib.apt_tile_plan(l, m, taps.ctypes.data, taps.size, C.byref(info), tt.ctypes.data, tt.size,
                             xs.ctypes.data, xs.size) == 0
    # [g][slice lane][iteration][16*halves] (+ padding per sub-table); chunk = iteration*slices + lane
    rl = 16 * info.halves
    rec = tt.reshape(info.groups, info.slices, info.slice_stride)[:, :, : info.iters * rl]
    rec = rec.reshape(info.groups, info.slices, info.iters, rl).transpose(0, 2, 1, 3)   # -> [g][it][lane][rl]
    rec = rec.reshape(info.groups, info.iters * info.slices, rl)                         # [g][chunk][rl]
    ta = rec[:, :, 0:16].reshape(info.groups, info.usteps, 4)                      # [g][u][r]
    tb = rec[:, :, 16:32].reshape(info.groups, info.usteps, 4) if info.halves == 2 else None
    return info, ta, tb, xs


@pytest.mark.parametrize("rate,work,l,m", [(48000, 12480, 13, 50), (48000, 16640, 26, 75), (96000, 12480, 13, 100),
                                           (48000, 20800, 13, 30), (96000, 16640, 13, 75)])
def test_tile_plan_geometry_reproduces_fast_resampling(na, rate, work, l, m):
    """The tiled kernel's host-built geometry (groups, window starts, the two half windows, zero-padded tap
    records), emulated with numpy, must give fast_resampling's outputs (dsp.rs:186-289) -- vs the oracle."""
    import math
    import oracle
    prof = {12480: "standard", 16640: "fast", 20800: "slow"}[work]
    st = na.Settings.profile(prof)
    f = na.filters.LowpassDcRemoval(na.Freq.hz(st.resample_cutout, rate), st.resample_atten,
                                    na.Freq.hz(st.resample_delta_freq, rate))
    f.resample(rate, rate * l)
    h = f.design()
    info, ta, tb, xs = _tile_plan(na, l, m, h)
    assert info is not None and info.usable
    R = 4 * info.halves
    assert info.p_out * m == info.p_in * l and info.p_in % 4 == 0 and info.p_out == R * info.groups
    assert len({(ks * (info.slice_stride // 4)) % 8 for ks in range(4)}) == 4       # tap bank-group skew
    if info.rows_per_copy == 2:
        starts = {((r >> 1) * (info.pair_pitch // 4) + (r & 1) * (info.p_in // 4)) % 8 for r in range(8)}
        assert info.pair_pitch >= info.p_in + info.row_len
    else:
        starts = {(r * (info.pair_pitch // 4)) % 8 for r in range(8)}
        assert info.pair_pitch >= info.row_len
    assert len(starts) == 8                                                        # 8 row lanes, 8 bank groups
    assert info.usteps == info.half_taps + info.shift == 16 * info.iters and info.shift % 16 == 0
    assert info.halves == 2 or info.shift == 0
    assert all(int(v) % 4 == 0 and int(v) + info.usteps <= info.row_len for v in xs)
    assert info.ctas_per_sm == 1 and info.smem_bytes <= 227 * 1024
    # the loop skips half B before `shift` and half A after `half_taps`: those taps must be zero
    assert not ta[:, info.half_taps:].any() and (tb is None or not tb[:, : info.shift].any())
    x = (np.random.default_rng(0).standard_normal(30000) * 1000).astype(np.float32)
    ref = oracle.fast_resampling(x, l, m, h)
    qt, tile_out = info.rows_per_tile, info.rows_per_tile * info.p_out
    xpad = np.concatenate([x.astype(np.float64), np.zeros(qt * info.p_in + info.row_len + 16)])
    out = np.zeros(ref.size + tile_out)
    for t in range(math.ceil(ref.size / tile_out)):
        xb = t * qt * info.p_in
        for q in range(qt):
            for g in range(info.groups):
                win = xpad[xb + q * info.p_in + int(xs[g]): xb + q * info.p_in + int(xs[g]) + info.usteps]
                k = t * tile_out + q * info.p_out + R * g
                out[k:k + 4] = win @ ta[g].astype(np.float64)
                if tb is not None:
                    out[k + 4:k + 8] = win @ tb[g].astype(np.float64)
    assert np.max(np.abs(out[:ref.size] - ref)) <= 1e-6 * np.max(np.abs(ref))


def test_tile_plan_rejects_shapes_it_cannot_hold(na):
    h = np.ones(14057, np.float32)
    info, _, _, _ = _tile_plan(na, 832, 735, h)          # 11025 Hz: 208 groups > 13 warps -> generic kernel
    assert info is None


def _ut_plan(na, l, m, taps):
    import ctypes as C
    from noaa_apt_b200._lib import CUtInfo
    lib = na._lib.load() if hasattr(na, "_lib") else None
    info = CUtInfo()
    taps = np.ascontiguousarray(taps, dtype=np.float32)
    assert lib.apt_ut_plan(l, m, taps.ctypes.data, taps.size, C.byref(info), None, 0) == 0
    if not info.usable:
        return info, None
    stream = np.zeros(4 * info.nvec, dtype=np.float32)
    assert lib.apt_ut_plan(l, m, taps.ctypes.data, taps.size, C.byref(info), stream.ctypes.data, stream.size) == 0
    return info, stream


@pytest.mark.parametrize("rate,work,l,m", [(48000, 12480, 13, 50), (96000, 12480, 13, 100), (192000, 12480, 13, 200),
                                           (48000, 20800, 13, 30), (96000, 20800, 13, 60)])
def test_uniform_tap_plan_reproduces_fast_resampling(na, rate, work, l, m):
    """The uniform-tap kernel's host-built plan (roles, chunk ranges per pair, the tap stream in consumption order),
    emulated with numpy exactly the way the kernel's loop walks it, must give fast_resampling's outputs
    (dsp.rs:186-289) -- checked against the oracle."""
    import oracle
    prof = {12480: "standard", 20800: "slow"}[work]
    st = na.Settings.profile(prof)
    f = na.filters.LowpassDcRemoval(na.Freq.hz(st.resample_cutout, rate), st.resample_atten,
                                    na.Freq.hz(st.resample_delta_freq, rate))
    f.resample(rate, rate * l)
    h = f.design()
    info, stream = _ut_plan(na, l, m, h)
    assert info.usable and info.l == l and info.m == m and info.np == (l + 1) // 2
    cs, ce, CH = list(info.cs), list(info.ce), info.chunk_len
    assert CH % 4 == 0
    npa = (info.np + 1) // 2
    assert info.rows_per_block == 32 * info.q and info.back % 4 == 0 and info.slot_floats % 4 == 0
    assert info.slot_floats >= info.back + (info.rows_per_block - 1) * m + CH * info.chunks
    assert info.slot_floats >= info.rows_per_block * l                   # outputs are staged in the slot
    assert info.warps % 2 == 0 and info.nslot > info.warps // 2 and info.smem_bytes <= 227 * 1024
    assert info.vec == (4 if m % 4 == 0 else 2 if m % 2 == 0 else 1)
    assert info.nvec * 16 + 64 <= 32764 - 256                            # fits the kernel-parameter space
    # the kernel's loop: role by role, segment by segment, 2*CH taps per (chunk, active pair)
    taps4 = stream.reshape(-1, CH, 2).astype(np.float64)                # [record][sample in chunk][output in pair]
    T = np.zeros((CH * info.chunks, 2 * info.np))                         # rebuilt tap matrix T[u][r]
    seen = np.zeros_like(T, dtype=bool)
    rec = 0
    for pb, npr in ((0, npa), (npa, info.np - npa)):
        assert rec * CH // 2 == (0 if pb == 0 else info.stream_b)   # stream_b counts float4, a record is CH/2 of them
        segs = [(cs[pb + a - 1], cs[pb + a], pb, pb + a) for a in range(1, npr)]
        segs.append((cs[pb + npr - 1], ce[pb], pb, pb + npr))
        segs += [(ce[pb + a - 1], ce[pb + a], pb + a, pb + npr) for a in range(1, npr)]
        for c0, c1, p0, p1 in segs:
            assert c0 <= c1
            for c in range(c0, c1):
                for p in range(p0, p1):
                    assert not seen[CH * c, 2 * p]
                    T[CH * c:CH * c + CH, 2 * p:2 * p + 2] = taps4[rec]
                    seen[CH * c:CH * c + CH, 2 * p:2 * p + 2] = True
                    rec += 1
    assert rec * CH // 2 == info.nvec
    # every tap of the filter appears exactly once: T[u][r] = h[u*l - r*m]
    off2 = 2 * ((len(h) - 1) // 2)
    hh = np.asarray(h, dtype=np.float64)
    for r in range(l):
        u = np.arange(CH * info.chunks)
        idx = u * l - r * m
        ok = (idx >= 0) & (idx <= off2)
        want = np.where(ok, hh[np.clip(idx, 0, off2)], 0.0)
        assert np.array_equal(T[:, r], want)
        assert ((r * m + off2) // l) < CH * info.chunks                   # the row window covers the last tap
    assert not T[:, l:].any()                                             # the padding output of an odd L
    # halo output (r = l-1 of the row in front of a block)
    assert info.halo_u0 == -(-((l - 1) * m) // l) and info.back >= m - info.halo_u0
    assert info.halo_u0 + info.halo_n - 1 == ((l - 1) * m + off2) // l
    x = (np.random.default_rng(0).standard_normal(20000) * 1000).astype(np.float32)
    ref = oracle.fast_resampling(x, l, m, h)
    rows = -(-ref.size // l)
    xpad = np.concatenate([x.astype(np.float64), np.zeros(rows * m + CH * info.chunks)])
    win = np.lib.stride_tricks.sliding_window_view(xpad, CH * info.chunks)[::m][:rows]    # [row][u]
    out = (win @ T[:, :l]).reshape(-1)[:ref.size]
    # the oracle sums up to 428 taps per output in f32; this emulation sums in f64 (the taps were checked exactly above)
    assert np.max(np.abs(out - ref)) <= 5e-6 * np.max(np.abs(ref))


def test_uniform_tap_plan_rejects_other_ratios(na):
    h = np.ones(959, dtype=np.float32)
    for l, m in ((832, 735), (26, 75), (208, 735)):
        info, _ = _ut_plan(na, l, m, h)
        assert not info.usable


@pytest.mark.parametrize("in_rate,work_rate,atten,dfreq", [(11025, 12480, 30.0, 1000.0), (22050, 12480, 30.0, 1000.0),
                                                           (44100, 12480, 30.0, 1000.0), (11025, 20800, 40.0, 500.0)])
def test_phase_major_plan_reproduces_fast_resampling(in_rate, work_rate, atten, dfreq):
    """Host logic of kernels_ph.cuh: y[l*q + r] = sum_i table[r/4][i][r%4] * x[m*q + xs[r/4] - 4 + i] must be
    fast_resampling (dsp.rs:186-289) -- emulated in numpy (float64 accumulation) against the oracle."""
    import math
    import oracle
    from noaa_apt_b200 import _lib
    lib = _lib.load()
    g = math.gcd(in_rate, work_rate)
    l, m = work_rate // g, in_rate // g
    cut, dw = oracle.freq_hz(4800.0, in_rate), oracle.freq_hz(dfreq, in_rate)
    ratio = np.float32(in_rate * l) / np.float32(in_rate)
    taps = oracle.design(oracle.FILTER_LOWPASS_DC, float(np.float32(cut) / ratio), atten, float(np.float32(dw) / ratio))
    info = _lib.CPhInfo()
    assert lib.apt_ph_plan(l, m, taps.ctypes.data, taps.size, C.byref(info), None, 0, None, 0) == 0
    if in_rate == 11025 and work_rate == 20800:
        # slow profile at 11025 Hz: 40 891 taps -> 50 per output, table too large for shared memory: generic kernel
        assert info.usable in (0, 1)
        if not info.usable:
            return
    assert info.usable == 1 and info.l == l and info.m == m and info.jpad in (24, 44, 84) and info.pitch % 32 == 4
    table = np.zeros(l * info.jpad, np.float32)
    xs = np.zeros(l // 4, np.uint16)
    assert lib.apt_ph_plan(l, m, taps.ctypes.data, taps.size, C.byref(info), table.ctypes.data, table.size, xs.ctypes.data, xs.size) == 0
    assert (xs % 4 == 0).all()                              # 16-byte aligned windows
    table = table.reshape(l // 4, info.jpad, 4)
    rng = np.random.default_rng(7)
    x = (rng.standard_normal(m * 40 + 123) * 1000).astype(np.float32)
    ref = oracle.fast_resampling(x, l, m, taps)
    xp = np.concatenate([np.zeros(4), x.astype(np.float64), np.zeros(m + 200)])   # xp[4 + i] = x[i]
    k = np.arange(ref.size)
    q, r = k // l, k % l
    idx = (q * m + xs[r // 4].astype(np.int64))[:, None] + np.arange(info.jpad)[None, :]   # row index: signal index + 4
    got = np.sum(table[r // 4, :, r % 4].astype(np.float64) * xp[idx], axis=1)
    assert np.max(np.abs(got - ref)) <= 1e-5 * np.max(np.abs(ref))


def test_wav_reader_and_writer_roundtrip(tmp_path):
    """wav.rs:11-56 restated without hound: 16-bit mono, stereo (channel 0 kept), 8-bit (offset 128), 32-bit float; the
    16-bit writer of resample.rs:53-66.  No GPU involved."""
    import struct
    import wave
    from noaa_apt_b200 import wav
    rng = np.random.default_rng(5)
    pcm = rng.integers(-32768, 32767, 5000, dtype=np.int16)
    p = str(tmp_path / "mono16.wav")
    with wave.open(p, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(11025); w.writeframes(pcm.tobytes())
    assert wav.info(p) == {"sample_rate": 11025, "channels": 1, "bits_per_sample": 16, "is_float": False, "frames": 5000}
    x, rate = wav.load_wav(p)
    assert rate == 11025 and np.array_equal(x, pcm.astype(np.float32))          # `as f32`: raw values
    raw, _ = wav.load_wav_pcm16(p)
    assert np.array_equal(raw, pcm)
    # stereo: interleaved, channel 0 kept
    st = np.stack([pcm, -pcm], axis=1).reshape(-1)
    p2 = str(tmp_path / "stereo16.wav")
    with wave.open(p2, "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(48000); w.writeframes(st.tobytes())
    x2, rate2 = wav.load_wav(p2)
    assert rate2 == 48000 and np.array_equal(x2, pcm.astype(np.float32))
    # 8-bit unsigned
    u8 = rng.integers(0, 255, 1000, dtype=np.uint8)
    p3 = str(tmp_path / "mono8.wav")
    with wave.open(p3, "wb") as w:
        w.setnchannels(1); w.setsampwidth(1); w.setframerate(8000); w.writeframes(u8.tobytes())
    x3, _ = wav.load_wav(p3)
    assert np.array_equal(x3, u8.astype(np.float32) - 128)
    # 32-bit float (format tag 3), written by hand, with an odd-sized LIST chunk in front of the data
    f32 = rng.standard_normal(777).astype(np.float32)
    p4 = str(tmp_path / "f32.wav")
    body = (b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 3, 1, 96000, 96000 * 4, 4, 32) + b"LIST" + struct.pack("<I", 3) + b"abc\0" +
            b"data" + struct.pack("<I", f32.nbytes) + f32.tobytes())
    open(p4, "wb").write(b"RIFF" + struct.pack("<I", len(body)) + body)
    x4, rate4 = wav.load_wav(p4)
    assert rate4 == 96000 and np.array_equal(x4, f32)
    # writer
    p5 = str(tmp_path / "out.wav")
    wav.write_wav_i16(p5, pcm, 12480)
    with wave.open(p5) as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, 12480, 5000)
        assert np.array_equal(np.frombuffer(w.readframes(5000), dtype="<i2"), pcm)
    import noaa_apt_b200 as na
    with pytest.raises(na.err.Io):
        wav.load_wav(str(tmp_path / "missing.wav"))
