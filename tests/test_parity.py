"""Parity of the product (through the C ABI) against the CPU oracle.

Every test runs on two backends (tests/common.py): "emu" on CPU (-m "not gpu") and
"cuda" on the B200 (-m gpu).  Tolerance: 1e-9 RMS per block (BASELINE.json
north_star) for floating-point units; bit-exact where the arithmetic is exact
(pass-through, noise gate, tremolo, integer-PCM round trip of pointwise units).
"""
import numpy as np
import pytest

import common
from common import gdg, orc

BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


@pytest.fixture(params=BACKENDS)
def lib(request):
    return common.backend_lib(request.param)


UNIT_CASES = [
    # (unit, numeric, discrete)
    ("signal_generator", {}, {}),
    ("signal_generator", {"signal_frequency": 997, "input_gain": -6}, {"signal_type": "triangle"}),
    ("signal_generator", {"signal_frequency": 1234}, {"signal_type": "square"}),
    ("signal_generator", {"signal_frequency": 55, "signal_gain": -12}, {"signal_type": "sawtooth"}),
    ("signal_generator", {"signal_amplitude": 50}, {"signal_type": "noise"}),
    ("noise_gate", {}, {}),
    ("noise_gate", {"threshold_open": -12, "threshold_close": -30, "hold_time": 3}, {}),
    ("noise_gate", {"threshold_open": -40, "threshold_close": -20}, {}),          # open < close: copy
    ("noise_gate", {"hold_time": 0}, {}),
    ("bandpass", {}, {}),
    ("bandpass", {"frequency_1": 5000, "frequency_2": 200}, {"filter_order": "8"}),
    ("auto_wah", {}, {}),
    ("auto_wah", {"level_1": -5, "level_2": -50}, {"follow": "envelope"}),
    ("auto_yoy", {}, {}),
    ("auto_yoy", {"depth": 60}, {"follow": "envelope"}),
    ("compressor", {}, {}),
    ("compressor", {"gain_limit": 12, "target_level": -6}, {"follow": "envelope"}),
    ("octaver", {}, {}),
    ("octaver", {"level_hysteresis": -6, "level_octave_up": -10}, {"follow": "envelope"}),
    ("excess", {"gain": 18}, {}),
    ("excess", {"gain": 18}, {"oversampling": "2"}),
    ("excess", {"gain": 25, "level": -3}, {"oversampling": "4"}),
    ("fuzz", {"gain": 12}, {}),
    ("fuzz", {"gain": 12, "bias": -30, "fuzz": 70}, {"oversampling": "2", "follow": "envelope"}),
    ("fuzz", {"gain": 20, "boost": 10}, {"oversampling": "4"}),
    ("overdrive", {"gain": 20}, {}),
    ("overdrive", {"gain": 20, "drive": 60}, {"oversampling": "2", "valve": "ECC82 (12AU7)"}),
    ("overdrive", {"gain": 30, "boost": 12}, {"oversampling": "4"}),
    ("distortion", {"gain": 15}, {}),
    ("distortion", {"gain": 15}, {"oversampling": "2"}),
    ("distortion", {"gain": 25, "level": -6}, {"oversampling": "4"}),
    ("tone_stack", {}, {}),
    ("tone_stack", {"low": -12, "middle": 0, "presence": -20, "high": -1}, {}),
    ("chorus", {}, {}),
    ("chorus", {"depth": 35, "speed": 100}, {}),
    ("flanger", {}, {}),
    ("flanger", {"depth": 50, "speed": 77}, {}),
    ("phaser", {}, {}),
    ("phaser", {"phase": -90, "speed": 33}, {}),
    ("tremolo", {}, {}),
    ("tremolo", {"frequency": 37, "phase": 0}, {}),
    ("tremolo", {"frequency": 100, "phase": 100, "depth": -40}, {}),
    ("ring_modulator", {}, {}),
    ("ring_modulator", {"frequency": 7}, {}),
    ("delay", {}, {}),
    ("delay", {"delay_time": 0}, {}),
    ("delay", {"delay_time": 1000, "feedback": 0, "level": 0}, {}),
    ("reverb", {}, {}),
    ("reverb", {"mix": 100}, {}),
    ("cabinet", {}, {}),
]


def _silent_by_design(case):
    """parameter corners whose correct output is silence on a short stream: a gate that never holds or never opens (threshold 0 dB, input below full scale), a 100 % wet reverb
    before its first tap (192 ms), a power amp without a filter"""
    u, n, d = case
    return (u == "noise_gate" and (n.get("hold_time", 1) == 0 or n.get("threshold_open") == 0)) or (u == "reverb" and n.get("mix") == 100) or u == "power_amp"


def _id(case):
    u, n, d = case
    return u + ("-" + "-".join("%s=%s" % kv for kv in list(n.items()) + list(d.items())) if (n or d) else "")


@pytest.mark.parametrize("case", UNIT_CASES, ids=_id)
def test_unit_parity(lib, case):
    """One unit, 6 consecutive blocks (state carry), two block sizes and two sample rates."""
    for n, fs in ((8192, 192000), (1024, 96000), (64, 44100)):
        a, b = common.make_chain_pair(lib, [case])
        x = common.synth_fast(3, fs, 6 * n, seed=n)
        x *= 1.6                                            # drive the clamps and thresholds
        common.compare_stream(a, b, x, n, fs, label="%s n=%d fs=%d" % (_id(case), n, fs), expect_signal=not _silent_by_design(case))


def test_bit_exact_units(lib):
    """Units whose arithmetic involves no reassociation are bit-identical to the oracle."""
    fs, n = 96000, 2048
    for case in [("noise_gate", {}, {}), ("tremolo", {}, {}), ("distortion", {"gain": 15}, {}),
                 ("excess", {"gain": 18}, {}), ("delay", {}, {}), ("signal_generator", {}, {"signal_type": "noise"})]:
        a, b = common.make_chain_pair(lib, [case])
        x = common.synth_fast(1, fs, 4 * n)
        for k in range(0, len(x), n):
            ya, yb = a.process(x[k:k + n], fs), b.process(x[k:k + n], fs)
            assert np.array_equal(ya, yb), case


def test_empty_and_bypassed_chain_is_identity(lib):
    c = gdg.Chain(lib=lib)
    x = common.synth_fast(0, 48000, 4096)
    assert np.array_equal(c.process(x, 48000), x)
    i = c.append_unit("distortion")                          # appended units start bypassed (signal.go:74)
    assert c.get_bypass(i) is True
    assert np.array_equal(c.process(x, 48000), x)


def test_parameter_change_between_blocks(lib):
    """Setters take effect at the next block boundary; state survives the change."""
    fs, n = 96000, 4096
    a, b = common.make_chain_pair(lib, [("overdrive", {"gain": 10}, {"oversampling": "4"}), ("tone_stack", {}, {}),
                                        ("delay", {"delay_time": 20}, {})])
    x = common.synth_fast(5, fs, 8 * n)
    for blk in range(8):
        if blk == 2:
            for ch in (a, b):
                ch.set_numeric_value(0, "gain", 30)
                ch.set_numeric_value(1, "middle", -20)
        if blk == 4:
            for ch in (a, b):
                ch.set_discrete_value(0, "oversampling", "2")      # separate x2 oversampler state
                ch.set_numeric_value(2, "delay_time", 35)          # ring re-created (zeroed)
        if blk == 6:
            for ch in (a, b):
                ch.set_bypass(1, True)
                ch.move_down(0)                                    # delay before overdrive now? no: overdrive <-> tone_stack
        ya, yb = a.process(x[blk * n:(blk + 1) * n], fs), b.process(x[blk * n:(blk + 1) * n], fs)
        assert common.rms(ya - yb) <= common.TOL_RMS, blk


def test_sample_rate_change_rebuilds_state(lib):
    a, b = common.make_chain_pair(lib, [("chorus", {}, {}), ("reverb", {}, {})])
    for fs in (48000, 96000, 48000):
        x = common.synth_fast(2, fs, 3 * 2048)
        common.compare_stream(a, b, x, 2048, fs, label="fs=%d" % fs)


# ---- power amp -----------------------------------------------------------------------------------
@pytest.mark.parametrize("fs,n,order,ir", [
    (96000, 8192, "8192", "Guitar: American Vintage (Center)"),      # config 2: 3839 taps -> F = 8192
    (192000, 8192, "8192", "Guitar: American Vintage (Center)"),     # 7679 taps -> F = 16384
    (192000, 8192, "8192", "Guitar: Custom BP"),                     # 15359 taps reduced to 8192 (filter.Reduce)
    (48000, 1024, "1048576", "Guitar: British Vintage (Bright)"),    # short IR, many blocks per call
    (44100, 256, "64", "PA: HC 4-way (Sub)"),                        # Reduce to 64 taps
    (96000, 512, "4096", "Guitar: American Modern (Classic)"),       # n < B: single FFT, long tail
])
def test_poweramp_parity(lib, fs, n, order, ir):
    banks = common.make_banks(lib)
    names = banks[0].names()
    assert names == banks[1].names()
    if ir not in names:
        ir = [k for k in names if k.startswith(ir.split(":")[0])][0]
    spec = [("power_amp", {"level_1": -3}, {"filter_order": order, "filter_1": ir})]
    a, b = common.make_chain_pair(lib, spec, banks)
    x = common.synth_fast(4, fs, 6 * n) * 3.0                 # the reference clamps the convolution output
    common.compare_stream(a, b, x, n, fs, label="poweramp")


def test_poweramp_mix_of_filters_and_none(lib):
    banks = common.make_banks(lib, synthetic=[("synthetic-300", 300, 96000)])
    names = banks[0].names()
    spec = [("power_amp", {"level_1": -6, "level_2": -12, "level_3": 0},
             {"filter_order": "2048", "filter_1": names[0], "filter_2": names[20], "filter_3": "synthetic-300"})]
    a, b = common.make_chain_pair(lib, spec, banks)
    x = common.synth_fast(6, 96000, 5 * 4096)
    common.compare_stream(a, b, x, 4096, 96000, label="poweramp-mix")
    # all filters "- NONE -": zeros (filter.go:366-367); no bank at all: zeros (poweramp.go:210-214)
    for bk in (banks, None):
        a2, b2 = common.make_chain_pair(lib, [("power_amp", {}, {})], bk)
        y = b2.process(x[:4096], 96000)
        assert np.array_equal(y, np.zeros(4096)) and np.array_equal(a2.process(x[:4096], 96000), y)


@pytest.mark.parametrize("L", [256, 1024, 8192, 16384, 20000, 65536])
def test_ir_sweep_synthetic(lib, L):
    """BASELINE.json config 3: synthetic IRs of 256 .. 65536 taps at 192 kHz, block 8192.  Up to 8192 taps the
    whole FFT sits in shared memory; longer filters run as a uniformly partitioned convolution, which must still
    equal the reference's single 2*nextPow2(L) FFT per hop (filter.go:372-455)."""
    banks = common.make_banks(lib, with_fixture=False, synthetic=[("syn", L, 192000)])
    spec = [("power_amp", {}, {"filter_1": "syn"})]
    a, b = common.make_chain_pair(lib, spec, banks)
    x = common.synth_fast(7, 192000, 4 * 8192)
    common.compare_stream(a, b, x, 8192, 192000, label="sweep L=%d" % L)


@pytest.mark.parametrize("n", [64, 256, 1024, 4096])
def test_long_filter_at_jack_block_sizes(lib, n):
    """The reference's default filter_order (1048576: no reduction) leaves 7 of the 24 bank responses above 8192 taps at
    192 kHz (and one at 96 kHz); the reference convolves them at any JACK period with one 2*nextPow2(L) FFT per block
    (filter.go:372-455, doc/jack-setup.txt:15-16).  Here they run as a uniformly partitioned convolution whose partition is
    the block length: same linear convolution, same overlap-add, same clamp."""
    banks = common.make_banks(lib, synthetic=[("syn20000", 20000, 192000)])
    for name, fs in (("Guitar: Custom BP", 192000), ("syn20000", 192000), ("PA: HC 4-way (Sub)", 96000)):
        a, b = common.make_chain_pair(lib, [("power_amp", {"level_1": -3}, {"filter_1": name})], banks)
        blocks = max(6, 3 * 20000 // n // 4)
        x = common.synth_fast(5, fs, blocks * n) * 1.5
        common.compare_stream(a, b, x, n, fs, label="%s n=%d" % (name, n))


def test_long_filter_block_length_change_repartitions(lib):
    """DELIBERATE DIFFERENCE: a new block length re-partitions a filter longer than 8192 taps and restarts its
    overlap tail; the reference's single FFT keeps the tail.  After the change the product equals a FRESH reference filter
    fed from that block on; nothing is rejected."""
    fs = 192000
    banks = common.make_banks(lib, with_fixture=False, synthetic=[("syn", 20000, fs)])
    spec = [("power_amp", {}, {"filter_1": "syn"})]
    a, b = common.make_chain_pair(lib, spec, banks)
    x = common.synth_fast(2, fs, 4 * 8192)
    common.compare_stream(a, b, x[:2 * 8192], 8192, fs, label="8192")
    fresh = common.build_chain(orc.Chain(banks[0]), spec)
    common.compare_stream(fresh, b, x[2 * 8192:2 * 8192 + 8 * 1024], 1024, fs, label="after the switch to 1024")


def test_default_filter_order_with_long_bank_response(lib):
    """filter_order left at its default (1048576): "Guitar: Custom BP" is 15359 taps at 192 kHz -> 2 partitions."""
    banks = common.make_banks(lib)
    spec = [("power_amp", {"level_1": -3}, {"filter_1": "Guitar: Custom BP"})]
    a, b = common.make_chain_pair(lib, spec, banks)
    x = common.synth_fast(8, 192000, 5 * 8192) * 2.0
    common.compare_stream(a, b, x, 8192, 192000, label="custom-bp-default-order")


# ---- the canonical 12-stage chain ------------------------------------------------------------------
@pytest.mark.parametrize("fs", [96000, 192000])
def test_chain12_parity(lib, fs):
    banks = common.make_banks(lib)
    a, b = common.make_chain_pair(lib, common.CHAIN12, banks, ir_name="Guitar: American Vintage (Center)")
    n = 8192
    x = common.synth_channel(0, fs, 5 * n)
    worst = common.compare_stream(a, b, x, n, fs, label="CHAIN12")
    assert worst <= common.TOL_RMS


def test_chain_with_all_21_units(lib):
    banks = common.make_banks(lib)
    spec = [(u, {}, {}) for u in gdg.UNIT_TYPES]
    spec[19] = ("power_amp", {}, {"filter_order": "2048", "filter_1": "Guitar: British Modern (Center)"})
    a, b = common.make_chain_pair(lib, spec, banks)
    x = common.synth_fast(9, 96000, 4 * 2048)
    common.compare_stream(a, b, x, 2048, 96000, label="all21")


def test_process_batch_matches_per_channel(lib):
    """gdg_process_batch == the reference's per-channel goroutine fan-out (controller.go:2682-2705)."""
    fs, n, C = 96000, 2048, 5
    banks = common.make_banks(lib)
    specs = [common.CHAIN12, common.CHAIN12[:6], [("reverb", {}, {})], [], common.CHAIN12[6:]]
    pairs = [common.make_chain_pair(lib, s, banks, ir_name="Guitar: British Vintage (Dark)") for s in specs]
    x = np.stack([common.synth_fast(c, fs, 3 * n) for c in range(C)])
    for k in range(3):
        blk = x[:, k * n:(k + 1) * n]
        ref = np.stack([pairs[c][0].process(blk[c], fs) for c in range(C)])
        out = gdg.process_batch([p[1] for p in pairs], blk, fs)
        for c in range(C):
            assert common.rms(ref[c] - out[c]) <= common.TOL_RMS, (k, c)


def test_pcm16_round_trip_bit_exact(lib):
    """BASELINE.json config 1 shape: PCM16 in -> overdrive -> tone_stack -> PCM16 out, compared as integers."""
    fs, n = 48000, 8192
    rng = np.random.default_rng(3)
    pcm = (rng.standard_normal(4 * n) * 6000).clip(-32768, 32767).astype(np.int16)
    x = orc.pcm16_decode(pcm)
    spec = [("overdrive", {"gain": 12}, {}), ("tone_stack", {}, {})]
    a, b = common.make_chain_pair(lib, spec)
    ya = np.concatenate([a.process(x[k:k + n], fs) for k in range(0, len(x), n)])
    yb = np.concatenate([b.process(x[k:k + n], fs) for k in range(0, len(x), n)])
    pa, pb = orc.pcm16_encode(ya), orc.pcm16_encode(yb)
    # identical up to the truncation boundary effect of a last-ulp difference: allow <= 1 LSB on
    # a vanishing fraction of samples, and require exact equality for the overwhelming majority
    diff = np.abs(pa.astype(np.int32) - pb.astype(np.int32))
    assert diff.max() <= 1 and np.count_nonzero(diff) <= len(diff) // 10000
    # a distortion-only chain has no transcendental in its path: exactly bit-identical
    a2, b2 = common.make_chain_pair(lib, [("distortion", {"gain": 20}, {})])
    assert np.array_equal(orc.pcm16_encode(a2.process(x[:n], fs)), orc.pcm16_encode(b2.process(x[:n], fs)))


# ---- spatializer / tuner / controller ----------------------------------------------------------------
def _spat_pair(lib, C):
    a, b = orc.Spatializer(C), gdg.Spatializer(C, lib=lib)
    for c in range(C):
        for s in (a, b):
            s.set_azimuth(c, -70.0 + 20.0 * c)
            s.set_distance(c, 1.0 + 0.25 * c)
            s.set_level(c, 1.0 if c % 3 else 0.5)
    return a, b


@pytest.mark.parametrize("fs,n", [(192000, 8192), (96000, 1024), (44100, 16)])
def test_spatializer_parity(lib, fs, n):
    C = 8
    a, b = _spat_pair(lib, C)
    a.set_sample_rate(fs); b.set_sample_rate(fs)
    x = np.stack([common.synth_fast(c, fs, 4 * n) for c in range(C)])
    aux = 0.1 * common.synth_fast(99, fs, 4 * n)
    for k in range(4):
        blk = x[:, k * n:(k + 1) * n]
        ax = aux[k * n:(k + 1) * n] if k % 2 else None
        ya, yb = a.process(blk, ax), b.process(blk, ax)
        # same accumulation order (channels ascending, then aux) on one GPU: bit-exact up to FMA contraction
        assert common.rms(ya - yb) <= 1e-12, k


def test_spatializer_many_channels_group_partials(lib):
    """More than 32 channels: 32-channel group partial sums combined in ascending order (not the strictly
    sequential order of spatializer.go:170-308): differences are rounding-level."""
    C, fs, n = 70, 192000, 2048
    a, b = orc.Spatializer(C), gdg.Spatializer(C, lib=lib)
    for c in range(C):
        for s_ in (a, b):
            s_.set_azimuth(c, -80.0 + 160.0 * c / (C - 1)); s_.set_distance(c, 0.3 + 0.1 * (c % 17)); s_.set_level(c, 0.2 + 0.01 * c)
    a.set_sample_rate(fs); b.set_sample_rate(fs)
    x = np.stack([common.synth_fast(c, fs, 3 * n) for c in range(C)])
    for k in range(3):
        ya, yb = a.process(x[:, k * n:(k + 1) * n]), b.process(x[:, k * n:(k + 1) * n])
        assert common.rms(ya - yb) <= 1e-12 and np.abs(ya).max() > 1.0


def test_spatializer_default_position_is_plain_sum(lib):
    s = gdg.Spatializer(3, lib=lib)
    x = np.stack([common.synth_fast(c, 96000, 512) for c in range(3)])
    y = s.process(x)
    assert np.allclose(y[0], x.sum(axis=0), atol=1e-15) and np.array_equal(y[0], y[1])
    assert s.get_input_count() == 3 and s.get_output_count() == 2
    with pytest.raises(gdg.GdgError, match=r"within \[0, 10\]"):
        s.set_distance(0, 11.0)
    with pytest.raises(gdg.GdgError, match="Only 3 channels exist"):
        s.set_azimuth(3, 0.0)


@pytest.mark.parametrize("note", ["D2", "A2", "D3", "G3", "H3", "E4"])
def test_tuner_matches_reference_recordings(lib, note):
    """tuner/tuner_test.go through the product: note exact, |cents| <= 5, and equal to the oracle."""
    import os
    z = np.load(os.path.join(common.GOLD, "tuner_tail_%s.npz" % note))
    t, o = gdg.Tuner(lib=lib), orc.Tuner()
    # feed in uneven blocks to exercise the ring wrap-around (circular.go:33-71)
    s, fs = z["samples"], int(z["sample_rate"])
    pos = 0
    for blk in (8192, 5000, 96000 - 13192 - 7, 7):
        t.process(s[pos:pos + blk], fs); o.process(s[pos:pos + blk], fs); pos += blk
    r, ro = t.analyze(), o.analyze()
    assert r["note"] == note and -5 <= r["cents"] <= 5 and np.isfinite(r["frequency"])
    assert r["note"] == ro["note"] and r["cents"] == ro["cents"]
    assert abs(r["frequency"] - ro["frequency"]) <= 1e-9 * ro["frequency"]


def test_controller_process_config4(lib):
    """BASELINE.json config 4 in miniature: 8-ch CHAIN12, 4x OS, 192 kHz, tuner + spatializer on."""
    fs, n, C = 192000, 8192, 8
    banks = common.make_banks(lib)
    pairs = [common.make_chain_pair(lib, common.CHAIN12, banks, ir_name="Guitar: American Vintage (Center)") for _ in range(C)]
    sa, sb = _spat_pair(lib, C)
    sa.set_sample_rate(fs); sb.set_sample_rate(fs)
    ta, tb = orc.Tuner(), gdg.Tuner(lib=lib)
    x = np.stack([common.synth_fast(c, fs, 2 * n) for c in range(C)])
    for k in range(2):
        blk = x[:, k * n:(k + 1) * n]
        ya = orc.controller_process([p[0] for p in pairs], ta, 0, sa, blk, fs)
        yb = gdg.controller_process([p[1] for p in pairs], tb, 0, sb, blk, fs)
        assert ya.shape == yb.shape == (C + 2, n)
        for c in range(C + 2):
            assert common.rms(ya[c] - yb[c]) <= common.TOL_RMS, (k, c)
    assert ta.analyze()["note"] == tb.analyze()["note"]


@pytest.mark.parametrize("unit", [u for u in gdg.UNIT_TYPES if u != "power_amp"])
def test_parameter_extremes(lib, unit):
    """Each numeric parameter at its minimum and at its maximum (others default), each discrete parameter at each of its
    values: the ranges the reference's setters accept (effects.go:287-320) all reach the kernels."""
    fs, n = 96000, 2048
    probe = gdg.Chain(lib=lib)
    i = probe.append_unit(unit)
    params = probe.parameters(i)
    x = common.synth_fast(11, fs, 3 * n) * 1.2
    variants = []
    for p in params:
        if p["DiscreteValues"]:
            variants += [({}, {p["Name"]: v}) for v in p["DiscreteValues"]]
        else:
            variants += [({p["Name"]: p["Minimum"]}, {}), ({p["Name"]: p["Maximum"]}, {})]
    assert variants
    for numeric, discrete in variants:
        a, b = common.make_chain_pair(lib, [(unit, numeric, discrete)])
        common.compare_stream(a, b, x, n, fs, label="%s %s %s" % (unit, numeric, discrete), expect_signal=not _silent_by_design((unit, numeric, discrete)))
