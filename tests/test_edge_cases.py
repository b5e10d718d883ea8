"""Edge cases of the hot path on both backends: tiny blocks, every supported sample rate, long streams,
chain edits with live state, state isolation between chains, scattered host buffers."""
import ctypes as C

import numpy as np
import pytest

import common
from common import gdg, orc

BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


@pytest.fixture(params=BACKENDS)
def lib(request):
    return common.backend_lib(request.param)


MIXED = [("compressor", {}, {}), ("overdrive", {"gain": 18}, {"oversampling": "4"}), ("fuzz", {"gain": 10}, {"oversampling": "2"}),
         ("tone_stack", {}, {}), ("chorus", {}, {}), ("delay", {"delay_time": 3}, {}), ("reverb", {}, {}),
         ("power_amp", {}, {"filter_order": "512", "filter_1": None}), ("cabinet", {}, {})]


@pytest.mark.parametrize("n", [1, 2, 4, 8, 16, 32, 128, 512])
def test_tiny_and_small_blocks(lib, n):
    """Block lengths below the CTA width (JACK periods go down to 64; the oracle accepts any power of two)."""
    banks = common.make_banks(lib)
    a, b = common.make_chain_pair(lib, MIXED, banks, ir_name="Guitar: British Vintage (Neutral)")
    fs = 48000
    x = common.synth_fast(2, fs, 24 * n + 0) * 1.5
    common.compare_stream(a, b, x, n, fs, label="n=%d" % n)


@pytest.mark.parametrize("fs", [22050, 32000, 44100, 48000, 88200, 96000, 192000])
def test_every_supported_sample_rate(lib, fs):
    """filter.go:25-33: the seven rates the IR bank is resampled to."""
    banks = common.make_banks(lib)
    a, b = common.make_chain_pair(lib, MIXED, banks, ir_name="PA: ISP 2-way (Low)")
    x = common.synth_fast(4, fs, 4 * 2048)
    common.compare_stream(a, b, x, 2048, fs, label="fs=%d" % fs)


def test_long_stream_does_not_drift(lib):
    """60 consecutive blocks: the parallel scans must not accumulate error against the sequential reference."""
    fs, n = 96000, 1024
    banks = common.make_banks(lib)
    a, b = common.make_chain_pair(lib, common.CHAIN12, banks, ir_name="Guitar: American Vintage (Center)")
    x = common.synth_fast(6, fs, 60 * n)
    worst = common.compare_stream(a, b, x, n, fs, label="drift")
    assert worst <= common.TOL_RMS


def test_chain_edits_keep_unit_state(lib):
    """RemoveUnit / MoveUp / MoveDown / SetBypass move live units around: their DSP state travels with them
    (signal.go:91-162 operate on the slot slice; the unit objects are untouched)."""
    fs, n = 96000, 2048
    spec = [("delay", {"delay_time": 10}, {}), ("reverb", {}, {}), ("tone_stack", {}, {}), ("overdrive", {"gain": 12}, {"oversampling": "2"})]
    a, b = common.make_chain_pair(lib, spec)
    x = common.synth_fast(7, fs, 12 * n)
    for blk in range(12):
        for ch in (a, b):
            if blk == 2: ch.move_up(3)              # overdrive before tone_stack
            if blk == 4: ch.set_bypass(1, True)     # reverb off (keeps its tails for later)
            if blk == 6: ch.remove_unit(0)          # delay gone
            if blk == 7: ch.set_bypass(0, False)    # reverb back on with its old state
            if blk == 9: ch.move_down(0)
            if blk == 10:
                i = ch.append_unit("cabinet"); ch.set_bypass(i, False)
        ya, yb = a.process(x[blk * n:(blk + 1) * n], fs), b.process(x[blk * n:(blk + 1) * n], fs)
        assert common.rms(ya - yb) <= common.TOL_RMS, blk
    assert [a.unit_type(i) for i in range(a.length())] == [b.unit_type(i) for i in range(b.length())]


def test_chains_do_not_share_state(lib):
    """Two chains over one bank, processed alternately and in one batch, behave like two independent oracle chains."""
    fs, n = 96000, 2048
    banks = common.make_banks(lib)
    pairs = [common.make_chain_pair(lib, common.CHAIN12, banks, ir_name=nm) for nm in ("Guitar: Custom BP", "PA: HC 4-way (Mid)")]
    x = [common.synth_fast(c, fs, 6 * n) for c in range(2)]
    for k in range(6):
        if k % 2:      # alternately: one batched call ...
            out = gdg.process_batch([p[1] for p in pairs], np.stack([x[c][k * n:(k + 1) * n] for c in range(2)]), fs)
        else:          # ... or two single-channel calls in swapped order
            out = [None, None]
            for c in (1, 0):
                out[c] = pairs[c][1].process(x[c][k * n:(k + 1) * n], fs)
        for c in range(2):
            ref = pairs[c][0].process(x[c][k * n:(k + 1) * n], fs)
            assert common.rms(ref - out[c]) <= common.TOL_RMS, (k, c)


def test_edit_between_single_chain_and_batch_calls(lib):
    """A chain driven through gdg_chain_process (the JACK path) AND through gdg_process_batch: an edit that changes its
    segment structure (the oversampled unit leaves / rejoins its own kernel's run) must reach BOTH launch-table caches,
    whichever entry point sees the edited chain first.  (The tables used to share one 'descriptors re-made' flag: the
    single-chain call consumed it and the batch kept launching the old segment layout.)"""
    fs, n = 48000, 256
    spec = [("overdrive", {"gain": 10}, {"oversampling": "4"}), ("cabinet", {}, {})]
    pairs = [common.make_chain_pair(lib, spec) for _ in range(2)]
    x = [common.synth_fast(c, fs, 9 * n) for c in range(2)]

    def block(k, batched):
        blk = [x[c][k * n:(k + 1) * n] for c in range(2)]
        out = gdg.process_batch([p[1] for p in pairs], np.stack(blk), fs) if batched else [p[1].process(b, fs) for p, b in zip(pairs, blk)]
        for c in range(2):
            ref = pairs[c][0].process(blk[c], fs)
            assert common.rms(ref - out[c]) <= common.TOL_RMS, (k, c, batched)
    block(0, True)
    for k, (bypass, first) in enumerate([(True, False), (False, True), (True, True), (False, False)]):
        for ch in pairs[0]:
            ch.set_bypass(0, bypass)
        block(1 + 2 * k, first)                 # the edited chain is seen by one entry point ...
        block(2 + 2 * k, not first)             # ... and then by the other


def test_one_thread_per_channel_like_the_reference_fan_out(lib):
    """controller.go:2682-2705 / 3339-3341: the reference runs Chain.Process of every channel on its own goroutine, all at
    once.  Four threads, each driving ITS chain through gdg_chain_process for 24 blocks while the others do the same (and
    one more thread keeps editing a fifth chain): every channel equals its oracle chain."""
    import threading
    fs, n, blocks = 96000, 1024, 24
    banks = common.make_banks(lib)
    specs = [common.CHAIN12, MIXED, common.CHAIN12[:6], MIXED[2:]]
    pairs = [common.make_chain_pair(lib, sp, banks, ir_name="Guitar: American Vintage (Center)") for sp in specs]
    x = [common.synth_fast(c, fs, blocks * n) for c in range(len(pairs))]
    outs = [np.zeros(blocks * n) for _ in pairs]
    errors = []
    stop = threading.Event()

    def worker(c):
        try:
            for k in range(blocks):
                outs[c][k * n:(k + 1) * n] = pairs[c][1].process(x[c][k * n:(k + 1) * n], fs)
        except Exception as e:                               # noqa: BLE001
            errors.append((c, e))

    def editor():
        ch = common.build_chain(gdg.Chain(banks[1], lib=lib), [("power_amp", {}, {"filter_order": "2048", "filter_1": "Guitar: Custom BP"})])
        k = 0
        while not stop.is_set():
            ch.set_numeric_value(0, "level_1", -(k % 30)); ch.process(x[0][:n], fs)      # recompiles a filter every block
            k += 1
    ts = [threading.Thread(target=worker, args=(c,)) for c in range(len(pairs))] + [threading.Thread(target=editor)]
    for t in ts:
        t.start()
    for t in ts[:-1]:
        t.join()
    stop.set(); ts[-1].join()
    assert not errors, errors
    for c in range(len(pairs)):
        for k in range(blocks):
            ref = pairs[c][0].process(x[c][k * n:(k + 1) * n], fs)
            assert common.rms(ref - outs[c][k * n:(k + 1) * n]) <= common.TOL_RMS, (c, k)


def test_scattered_host_buffers(lib):
    """gdg_process_batch with non-contiguous, unpinned row pointers (the Go [][]float64 case)."""
    fs, n, Cn = 48000, 1024, 5
    chains_o = [common.build_chain(orc.Chain(), [("distortion", {"gain": 9 + c}, {})]) for c in range(Cn)]
    chains_p = [common.build_chain(gdg.Chain(lib=lib), [("distortion", {"gain": 9 + c}, {})]) for c in range(Cn)]
    rows_in = [np.ascontiguousarray(common.synth_fast(c, fs, n)) for c in range(Cn)]          # separate allocations
    rows_out = [np.zeros(n) for _ in range(Cn)]
    hs = (C.c_void_p * Cn)(*[c._h for c in chains_p])
    ip = (C.c_void_p * Cn)(*[r.ctypes.data for r in reversed(rows_in)])                       # deliberately not ascending
    op = (C.c_void_p * Cn)(*[r.ctypes.data for r in reversed(rows_out)])
    assert lib.gdg_process_batch(hs, Cn, ip, op, n, fs) == 0, lib.gdg_last_error()
    for c in range(Cn):
        ref = chains_o[c].process(rows_in[Cn - 1 - c], fs)
        assert np.array_equal(ref, rows_out[Cn - 1 - c])


def test_length_mismatch_and_null_arguments(lib):
    c = gdg.Chain(lib=lib)
    assert lib.gdg_chain_process(c._h, None, None, 64, 48000) < 0
    assert lib.gdg_process_batch(None, 0, None, None, 64, 48000) < 0
    assert lib.gdg_chain_set_bypass(None, 0, 1) < 0
    assert b"NULL" in lib.gdg_last_error()


SIGNALS = ["silence", "dc_full_scale", "overdriven_square", "impulse", "denormal", "nyquist", "step_after_silence"]


@pytest.mark.parametrize("kind", SIGNALS)
def test_extreme_finite_signals(lib, kind):
    """Signals at the edges of the float64 sample range (the reference clamps to [-1, 1] in most units and never checks
    its input): silence, DC at full scale, a square wave five times over full scale, one impulse, denormals, the Nyquist
    tone, and silence followed by a step - through every unit type in one chain and through CHAIN12's oversampled core."""
    fs, n, blocks = 96000, 2048, 5
    t = np.arange(blocks * n)
    x = {"silence": np.zeros(blocks * n),
         "dc_full_scale": np.ones(blocks * n),
         "overdriven_square": 5.0 * np.sign(np.sin(2 * np.pi * 997.0 * t / fs) + 1e-12),
         "impulse": (t == 3 * n // 2).astype(np.float64),
         "denormal": 1e-310 * np.sign(np.sin(2 * np.pi * 440.0 * t / fs) + 1e-12),
         "nyquist": np.where(t % 2 == 0, 1.0, -1.0) * 0.9,
         "step_after_silence": (t >= 2 * n + 17).astype(np.float64) * -0.75}[kind]
    banks = common.make_banks(lib)
    a, b = common.make_chain_pair(lib, MIXED, banks, ir_name="Guitar: British Vintage (Neutral)")
    common.compare_stream(a, b, x, n, fs, label=kind, expect_signal=kind not in ("silence", "denormal"))
    every = [(u, {}, ({"oversampling": "2"} if u in ("excess", "fuzz", "overdrive", "distortion") else {})) for u in gdg.UNIT_TYPES]
    every[19] = ("power_amp", {}, {"filter_order": "2048", "filter_1": "Guitar: British Modern (Center)"})
    a, b = common.make_chain_pair(lib, every, banks)
    common.compare_stream(a, b, x, n, fs, label=kind + "/all21", expect_signal=kind != "silence")


def test_chain_longer_than_one_fused_segment(lib):
    """The chain kernel fuses up to 24 active units per launch; signal.Chain has no length limit (signal.go:64-86), so a
    longer chain runs as consecutive segments with the block round-tripping through HBM between them - same result, and
    a batch may mix short and long chains; encoded payloads keep float64 rows between the segments."""
    fs, n = 96000, 2048
    unit_cycle = ["tone_stack", "delay", "compressor", "tremolo", "bandpass", "overdrive", "chorus", "noise_gate", "reverb", "cabinet"]
    long_spec = [(unit_cycle[k % len(unit_cycle)], ({"delay_time": 3 + k} if unit_cycle[k % len(unit_cycle)] == "delay" else {}), {}) for k in range(53)]
    short_spec = long_spec[:5]
    pairs = [common.make_chain_pair(lib, long_spec), common.make_chain_pair(lib, short_spec), common.make_chain_pair(lib, long_spec[:25])]
    x = np.stack([common.synth_fast(c, fs, 3 * n, 5) for c in range(3)])
    for k in range(3):
        blk = x[:, k * n:(k + 1) * n]
        ref = np.stack([p[0].process(blk[c], fs) for c, p in enumerate(pairs)])
        got = gdg.process_batch([p[1] for p in pairs], blk, fs)
        for c in range(3):
            assert common.rms(ref[c] - got[c]) <= common.TOL_RMS, (k, c)
    # a bypassed unit does not count towards the segment length
    for i in range(25, 53):
        pairs[0][0].set_bypass(i, True); pairs[0][1].set_bypass(i, True)
    blk = x[0, :n]
    assert common.rms(pairs[0][0].process(blk, fs) - pairs[0][1].process(blk, fs)) <= common.TOL_RMS
    # encoded rows through a multi-segment chain: decoded on the way into the first segment, encoded out of the last one
    a, b = common.make_chain_pair(lib, long_spec[:30])
    pcm = (np.random.default_rng(3).standard_normal(n) * 6000).clip(-32768, 32767).astype("<i2").tobytes()
    got = gdg.process_batch_encoded([b], [pcm], "PCM16", "IEEE64", n, fs)[0]
    assert common.rms(np.frombuffer(got, dtype=np.float64) - a.process(orc.wave_decode("PCM16", pcm), fs)) <= common.TOL_RMS


@pytest.mark.parametrize("unit", ["flanger", "phaser"])
@pytest.mark.parametrize("fs,n", [(22050, 32), (44100, 64), (88200, 128)])
def test_flanger_phaser_full_depth_history_edge(lib, unit, fs, n):
    """depth 100 at 22.05 / 44.1 / 88.2 kHz: the delay reaches 0.002 fs = history + a fraction, so the late tap of the first
    samples of a block asks for one sample BEFORE the history (the reference panics, flanger.go:80-84).  Product and oracle
    read the oldest history sample there (tests/test_oracle_crosscheck.py pins that choice against the literal restatement);
    600 blocks shorter than the history, speed 37: the LFO passes sin(phase) > 0.9955 at a block start many times."""
    a, b = common.make_chain_pair(lib, [(unit, {"depth": 100, "speed": 37}, {})])
    x = common.synth_fast(11, fs, 600 * n)
    common.compare_stream(a, b, x, n, fs, label="%s depth 100 @%d" % (unit, fs))


@pytest.mark.gpu
def test_oversampled_run_one_cta_per_channel_at_small_batches():
    """Small batches run an oversampled run as a cluster pipeline (k_os_pipe, one CTA per unit); big ones as one CTA per
    channel (k_os).  The rest of this suite uses small batches, so re-run the CHAIN12 parity cases in a process that has the
    cluster path switched off (GDG_NO_OS_PIPE): both shapes must match the oracle."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, GDG_NO_OS_PIPE="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_parity.py"), "-m", "gpu", "-q", "-x", "-k", "chain12 or all_21_units or process_batch_matches or parameter_change",
                        "-p", "no:cacheprovider"], env=env, cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert " passed" in r.stdout
