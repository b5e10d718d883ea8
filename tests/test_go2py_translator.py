"""The Go -> Python translator behind tests/golden/translated_reference_vectors.* (tests/golden/go2py/) on small Go programs
whose results follow from the Go specification - slices alias their backing array, struct values are copied by range / copy /
append while pointers are not, integer division truncates toward zero, sized integers wrap, `:=` in an inner block shadows,
iota repeats its expression, embedded structs promote fields and methods, switch does not fall through - and, where the
reference tree is present (the build container; not the GPU box), the translation of the reference's packages against the
reference's own unit-test vectors plus a bit-for-bit regeneration of a sample of the committed streams.  CPU only."""
import json
import os
import sys
import textwrap

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden", "go2py"))
import go2py       # noqa: E402
import goruntime   # noqa: E402
import load        # noqa: E402
from goruntime import G  # noqa: E402

REFERENCE = "/root/reference"
have_reference = os.path.isdir(os.path.join(REFERENCE, "effects"))


def run_go(tmp_path, src, pkg="p"):
    d = tmp_path / pkg
    d.mkdir(exist_ok=True)
    (d / "a.go").write_text(textwrap.dedent(src))
    return load.Loader(str(tmp_path)).load(pkg)


def test_slices_alias_and_copy_semantics(tmp_path):
    m = run_go(tmp_path, """
        package p

        func Alias() []float64 {
        	a := make([]float64, 6)
        	b := a[2:5]
        	b[0] = 7.0
        	c := b[1:]
        	c[1] = 9
        	return a
        }

        func CopyOverlap() []int {
        	a := []int{1, 2, 3, 4, 5}
        	copy(a[1:], a[:4])
        	return a
        }

        func AppendInPlace() ([]int, []int) {
        	a := make([]int, 3, 8)
        	b := append(a, 10)
        	c := append(a, 20)
        	return b, c
        }

        func DeleteIdiom() []string {
        	s := []string{"a", "b", "c", "d"}
        	s = append(s[:1], s[2:]...)
        	return s
        }

        func NilSlice() (int, bool) {
        	var s []float64
        	return len(s), s == nil
        }
        """)
    assert m.Alias().tolist() == [0.0, 0.0, 7.0, 0.0, 9.0, 0.0]
    assert m.CopyOverlap().tolist() == [1, 1, 2, 3, 4]                   # memmove semantics
    b, c = m.AppendInPlace()
    assert b.tolist() == [0, 0, 0, 20] and c.tolist() == [0, 0, 0, 20]   # both appends wrote the shared backing array
    assert m.DeleteIdiom().tolist() == ["a", "c", "d"]
    assert m.NilSlice() == (0, True)
    with pytest.raises(goruntime.GoPanic):
        m.Alias()[6]


def test_struct_values_pointers_and_embedding(tmp_path):
    m = run_go(tmp_path, """
        package p

        type base struct {
        	gain  float64
        	count int
        }

        func (this *base) bump() {
        	this.count++
        }

        type unit struct {
        	base
        	name string
        }

        type holder struct {
        	values   []unit
        	pointers []*unit
        }

        func RangeCopies() (int, int) {
        	h := holder{values: make([]unit, 2), pointers: make([]*unit, 2)}
        	for i := range h.pointers {
        		h.pointers[i] = &unit{name: "p"}
        	}
        	for _, u := range h.values {
        		u.bump()
        	}
        	for _, u := range h.pointers {
        		u.bump()
        	}
        	return h.values[0].count, h.pointers[0].count
        }

        func Promotion() (float64, int, string) {
        	u := unit{base: base{gain: 2}, name: "x"}
        	u.gain = 3
        	u.bump()
        	u.base.bump()
        	return u.gain, u.count, u.name
        }

        func IndexedFieldStore() float64 {
        	vs := make([]unit, 3)
        	vs[1].gain = 4.5
        	return vs[1].gain + vs[0].gain
        }
        """)
    assert m.RangeCopies() == (0, 1)                                     # range over values copies, over pointers does not
    assert m.Promotion() == (3.0, 2, "x")
    assert m.IndexedFieldStore() == 4.5


def test_integer_arithmetic(tmp_path):
    m = run_go(tmp_path, """
        package p

        func Div(a int, b int) (int, int) {
        	return a / b, a % b
        }

        func ConstDiv() (float64, float64, int) {
        	x := 1 / 2 * 3.0
        	var y float64 = 7
        	z := 7 / 2
        	return x, y / 2, z
        }

        func Wrap(seed uint64) (uint64, uint32, int32) {
        	a := 64979 * seed
        	b := uint32(4000000000) + uint32(1000000000)
        	c := int32(2147483647)
        	c++
        	return a, b, int32(uint32(c) + 1)
        }

        func Shifts(n int) (int, int, int) {
        	return 1 << 4, n >> 1, (n & 6) | 1
        }

        func FloatToInt(x float64) (int, uint32) {
        	return int(x), uint32(x)
        }
        """)
    assert m.Div(7, 2) == (3, 1) and m.Div(-7, 2) == (-3, -1) and m.Div(7, -2) == (-3, 1)
    assert m.ConstDiv() == (0.0, 3.5, 3)                                 # untyped 1 / 2 is integer division: 0
    a, b, c = m.Wrap(18446744073709551615)
    assert a == (64979 * 18446744073709551615) % 2 ** 64 and b == 705032704 and c == -2147483647
    assert m.Shifts(13) == (16, 6, 5)
    assert m.FloatToInt(-2.7) == (-2, (-2) & 0xFFFFFFFF) and m.FloatToInt(2.7) == (2, 2)
    with pytest.raises(goruntime.GoPanic):
        m.Div(1, 0)


def test_control_flow_scoping_and_constants(tmp_path):
    m = run_go(tmp_path, """
        package p

        import "math"

        const (
        	KIND_INVALID = iota - 1
        	KIND_A
        	KIND_B
        )

        const (
        	SCALE = 0.5 * math.Pi
        	N     = 4
        )

        var table = [...]float64{1, 2, 3, N}
        var counter int

        func Shadow(x int) (int, int) {
        	y := 1
        	if x > 0 {
        		y := 10
        		x := y + 1
        		_ = x
        	}
        	for i := 0; i < 3; i++ {
        		y := i
        		_ = y
        	}
        	return x, y
        }

        func Switch(k int) string {
        	r := "none"
        	switch k {
        	case KIND_A:
        		r = "a"
        		break
        	case KIND_B, 7:
        		r = "b"
        	default:
        		r = "d"
        	}
        	switch {
        	case k > 100:
        		r += "!"
        	}
        	return r
        }

        func Loops(n int) (int, int, float64) {
        	s := 0
        	for i := 0; i < n; i++ {
        		if i%2 == 0 {
        			continue
        		}
        		if i > 6 {
        			break
        		}
        		s += i
        	}
        	k := 0
        	for k < 5 {
        		k += 2
        	}
        	t := 0.0
        	for i, v := range table {
        		t += float64(i) * v
        	}
        	return s, k, t
        }

        func Global() int {
        	counter += 2
        	counter++
        	return counter
        }

        func Named() (q int, err error) {
        	q = 3
        	return
        }

        func Variadic(xs ...float64) float64 {
        	s := 0.0
        	for _, x := range xs {
        		s += x
        	}
        	return s
        }

        func Spread() float64 {
        	v := []float64{1, 2, 3.5}
        	return Variadic(v...) + Variadic() + Variadic(1, 2)
        }

        func Consts() (int, int, int, float64) {
        	return KIND_INVALID, KIND_A, KIND_B, SCALE
        }

        func Complex() (float64, float64) {
        	z := complex(1.0, 2.0) * complex(0.0, 1.0)
        	return real(z), imag(z)
        }

        func TypeSwitch(v interface{}) string {
        	switch v.(type) {
        	case float64:
        		return "f"
        	case string:
        		return "s"
        	}
        	return "?"
        }

        func CommaOk() (int, bool, int, bool) {
        	m := make(map[string]int)
        	m["a"] = 5
        	x, ok := m["a"]
        	y, ok2 := m["b"]
        	return x, ok, y, ok2
        }
        """)
    assert m.Shadow(1) == (1, 1)                                         # inner := never touches the outer variables
    assert [m.Switch(k) for k in (0, 1, 7, 5, 200)] == ["a", "b", "b", "d", "d!"]
    assert m.Loops(10) == (1 + 3 + 5, 6, 0 * 1 + 1 * 2 + 2 * 3 + 3 * 4.0)
    assert m.Global() == 3 and m.Global() == 6 and m.counter == 6
    assert m.Named() == (3, None)
    assert m.Spread() == 6.5 + 0.0 + 3.0
    assert m.Consts() == (-1, 0, 1, 0.5 * np.pi)
    assert m.Complex() == (-2.0, 1.0)
    assert m.TypeSwitch(1.5) == "f" and m.TypeSwitch("x") == "s" and m.TypeSwitch(3) == "?"
    assert m.CommaOk() == (5, True, None, False) or m.CommaOk() == (5, True, 0, False)


def test_float_literals_into_float_contexts(tmp_path):
    """an integer literal that Go types as float64 by context must not turn a later division into an integer division"""
    m = run_go(tmp_path, """
        package p

        type state struct {
        	level float64
        	buf   []float64
        }

        func half(x float64) float64 {
        	return x / 2
        }

        func Run() (float64, float64, float64, float64) {
        	s := state{level: 1, buf: make([]float64, 2)}
        	s.buf[0] = 3
        	var acc float64 = 0
        	acc = 1
        	return s.level / 2, s.buf[0] / 2, acc / 2, half(5)
        }

        func Ret() float64 {
        	return 7
        }

        func UseRet() float64 {
        	return Ret() / 2
        }
        """)
    assert m.Run() == (0.5, 1.5, 0.5, 2.5) and m.UseRet() == 3.5


def test_unsupported_constructs_are_rejected(tmp_path):
    """what the translator does not implement fails at translation time - it is never guessed at"""
    for k, body in enumerate(("go f()", "defer f()", "select {\n}", "h := func() {}\n_ = h", "goto done")):
        with pytest.raises(go2py.GoSyntaxError):
            run_go(tmp_path, "package p\n\nfunc f() {}\n\nfunc g() {\n\t%s\n}\n" % body.replace("\n", "\n\t"), pkg="q%d" % k)


# ---- with the reference tree (build container only) ----------------------------------------------------------------------
@pytest.mark.skipif(not have_reference, reason="needs the reference tree (build container)")
def test_translated_packages_reproduce_the_reference_unit_tests():
    import selfcheck
    os.environ["GO2PY_SELFCHECK_TUNER"] = "1"
    n = selfcheck.run(load.Loader(REFERENCE))
    assert n >= 72
    # every integer division the translated packages executed is an integer division in the Go source (index arithmetic)
    known = {"fft.go:254", "fft.go:573", "fft.go:580", "fft.go:774", "fft.go:893", "filter.go:374", "resample.go:164", "wave.go:486"}
    assert {s for s in G.intdiv_sites if not s.startswith("tuner.go") and not s.startswith("a.go")} <= known, sorted(G.intdiv_sites)


@pytest.mark.skipif(not have_reference, reason="needs the reference tree (build container)")
def test_committed_vectors_regenerate_bit_for_bit():
    """a sample of the committed streams, regenerated now from /root/reference: the fixtures are what the generator writes"""
    import make_reference_vectors as gen
    M = json.load(open(os.path.join(HERE, "golden", "translated_reference_vectors.json")))
    Z = np.load(os.path.join(HERE, "golden", "translated_reference_vectors.npz"))
    ref = gen.Ref(REFERENCE)
    picked = [c for c in M["cases"] if c["kind"] in ("chain", "power_amp", "varn", "ratechange")][:14] + M["cases"][0:260:13]
    for case in picked:
        ch = ref.chain([(u, nu, di) for u, nu, di in case["spec"]], case["ir"])
        x = case["amp"] * Z[case["input"]]
        ys, pos = [], 0
        for k, n in enumerate(case["ns"]):
            for op in case["edits"].get(str(k), ()):
                r = getattr(ch, op[0])(*op[1:])
                assert (r[-1] if isinstance(r, tuple) else r) is None
            ys.append(ref.process(ch, x[pos:pos + n], case["fss"][k]))
            pos += n
        assert np.array_equal(np.concatenate(ys), Z[case["id"]]), case["label"]


@pytest.mark.skipif(not have_reference, reason="needs the reference tree (build container)")
def test_random_chains_oracle_equals_executed_reference():
    """tests/golden/go2py/hunt.py on 80 fresh seeds (4400 were hunted in round 2 without a mismatch): random chains of 1 - 6 units,
    parameter extremes, any rate, varying block lengths - the C oracle against the executed reference at 1e-12 of the peak"""
    import hunt
    assert hunt.hunt(20000, 20080, REFERENCE, verbose=False) == []


@pytest.mark.skipif(not have_reference, reason="build container only: the vectors were generated with this machine's libm")
def test_oracle_is_bit_identical_to_the_executed_reference_here():
    """with the same libm under both (Python's math and the C oracle link glibc here) the oracle does not just stay within
    1e-12 of the executed reference - every one of the committed streams is reproduced bit for bit"""
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
    import common
    import test_reference_translated as T
    bank = common.orc.IRBank()
    common.fill_bank_from_fixture(bank, common.orc.pcm16_decode)
    different = [c["id"] for c in T.CASES
                 if not np.array_equal(T.run_case(common.build_chain(common.orc.Chain(bank), T.spec_of(c), c["ir"]), c), T.Z[c["id"]])]
    assert different == []


@pytest.mark.skipif(not have_reference, reason="needs the reference tree (build container)")
def test_edit_sequences_oracle_equals_executed_reference():
    """hunt.py --edits on 60 fresh seeds (3200 hunted in round 2 without a mismatch): a live chain edited between blocks - append /
    remove / move / bypass / parameter changes (valid and invalid: the error strings must coincide too), new block lengths and
    rates"""
    import hunt
    assert hunt.hunt_edits(40000, 40060, REFERENCE, verbose=False) == []


@pytest.mark.skipif(not have_reference, reason="needs the reference tree (build container)")
def test_random_chains_emulated_product_equals_executed_reference():
    """hunt.py --product on 60 fresh seeds (3300 hunted in round 2 without a mismatch): the kernel source (host emulation) against
    the executed reference at 1e-9 RMS per block"""
    import hunt
    assert hunt.hunt_product(60000, 60060, REFERENCE, verbose=False) == []


@pytest.mark.skipif(not have_reference, reason="needs the reference tree (build container)")
def test_spatializer_oracle_and_emulated_product_equal_executed_reference():
    """hunt.py --spat on 60 fresh seeds (3200 hunted in round 2 without a mismatch): 1 - 70 channels (beyond 32: the product's
    group-partial path), geometry at random including its extremes, aux now and then"""
    import hunt
    assert hunt.hunt_spatializer(70000, 70060, REFERENCE, verbose=False) == []


@pytest.mark.skipif(not have_reference, reason="needs the reference tree (build container)")
def test_compiled_filters_meters_resampler_equal_executed_reference():
    """hunt.py --compile / --meter / --resample on fresh seeds (1060 / 2150 / 600 hunted in round 2 without a mismatch): the power
    amp's compiled composite filter from the product's device-side IR compilation, the level meter's parallel formulation and
    resample.Time - emulated product and oracle against the executed reference"""
    import hunt
    import make_reference_vectors as gen
    ref = gen.Ref(REFERENCE)
    assert hunt.hunt_compile(90000, 90025, REFERENCE, ref=ref, verbose=False) == []
    assert hunt.hunt_meter(90000, 90080, REFERENCE, ref=ref, verbose=False) == []
    assert hunt.hunt_resample(90000, 90150, REFERENCE, ref=ref, verbose=False) == []


@pytest.mark.skipif(not have_reference, reason="needs the reference tree (build container)")
def test_tuner_equals_executed_reference_on_random_tones():
    """hunt.py --tuner on three fresh tones (300 hunted in round 2 without a mismatch)"""
    import hunt
    assert hunt.hunt_tuner(95000, 95003, REFERENCE, verbose=False) == []


@pytest.mark.skipif(not have_reference, reason="needs the reference tree (build container)")
def test_codecs_equal_executed_reference_byte_for_byte():
    """hunt.py --codecs on 300 fresh rows (6300 hunted in round 2 without a mismatch): wave.go's sample conversions, executed, vs
    the oracle's codecs and the product's fused load / store - all 36 format pairs"""
    import hunt
    assert hunt.hunt_codecs(96000, 96300, REFERENCE, verbose=False) == []


@pytest.mark.skipif(not have_reference, reason="needs the reference tree (build container)")
def test_controller_process_executed_equals_the_one_call_mirror():
    """hunt.py --controller on 40 fresh configurations (1540 hunted in round 2 without a mismatch): controller.process itself,
    executed - tuner feed, chain fan-out, level-meter ports, metronome row, metrMasterOutput, spatializer - against
    gdg_controller_process_aux (emulated product) and the oracle"""
    import hunt
    assert hunt.hunt_controller(97000, 97040, REFERENCE, verbose=False) == []


def test_channels_as_far_as_the_controller_needs_them(tmp_path):
    m = run_go(tmp_path, """
        package p

        type task struct {
        	x int
        }

        type ctl struct {
        	requests  chan task
        	responses chan bool
        }

        func (this *ctl) Run(n int) int {
        	for i := 0; i < n; i++ {
        		this.requests <- task{x: i}
        	}
        	got := 0
        	for i := 0; i < n; i++ {
        		if <-this.responses {
        			got++
        		}
        	}
        	return got
        }
        """)
    seen = []
    c = m.ctl.__new__(m.ctl)
    resp = G.Chan()
    c.__dict__["responses"] = resp
    c.__dict__["requests"] = G.Chan(0, lambda t: (seen.append(t.x), True)[1], resp)
    assert c.Run(4) == 4 and seen == [0, 1, 2, 3]
    with pytest.raises(goruntime.GoPanic):
        G.recv(resp)


@pytest.mark.skipif(not have_reference, reason="needs the reference tree (build container)")
def test_batch_block_lengths_three_ways_equal_executed_reference():
    """hunt.py --big on 8 fresh seeds: blocks of 4096 / 8192 samples at 96 / 192 kHz (the batch path's own shapes), chains of
    3 - 10 units, power amps on the long responses without reduction (the product's partitioned convolution) - executed
    reference vs oracle (1e-12) and vs the emulated product (1e-9 RMS)"""
    import hunt
    assert hunt.hunt_big(98000, 98008, REFERENCE, verbose=False) == []
