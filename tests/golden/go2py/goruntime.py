"""Run-time support for Go sources translated mechanically by go2py.py (test infrastructure - see README in this directory).

Go values map to Python values like this: float64 -> float, every integer type -> int (conversions mask to the width),
complex128 -> complex, string -> str, bool -> bool, nil -> None, struct -> instance of a GoStruct subclass (pointers to
structs are the same object), slice / array -> GoSlice (a view: offset + length + capacity into a shared backing list, so
sub-slices alias like in Go), map -> dict, error -> GoError, interface values -> the value itself.
Integer division and remainder truncate toward zero; float division by zero gives +-Inf / NaN; indexing outside a slice
raises (Go panics)."""
import math
import cmath
import struct as _struct


class GoPanic(Exception):
    pass


class GoError(object):
    def __init__(self, msg):
        self.msg = msg

    def Error(self):
        return self.msg

    def __repr__(self):
        return "GoError(%r)" % (self.msg,)


# ---------------------------------------------------------------- structs
class GoStruct(object):
    _fields = ()          # own field names, declaration order
    _embedded = ()        # names of embedded (anonymous) struct fields
    _fkinds = {}          # field name -> 'f' (float64), 'i' (integer), 'c' (complex) for literal coercion

    def _zero(self):      # overridden by generated classes: sets every field to its zero value
        pass

    def __init__(self, **kw):
        self._zero()
        for k, v in kw.items():
            setattr(self, k, v)

    @classmethod
    def _pos(cls, *vals):
        o = cls()
        for k, v in zip(cls._fields, vals):
            setattr(o, k, v)
        return o

    def __getattr__(self, name):                       # promoted fields and methods of embedded structs
        if name.startswith("__"):
            raise AttributeError(name)
        for e in type(self)._embedded:
            inner = self.__dict__.get(e)
            if inner is not None:
                try:
                    return getattr(inner, name)
                except AttributeError:
                    pass
        raise AttributeError("%s has no field or method %s" % (type(self).__name__, name))

    def __setattr__(self, name, value):
        cls = type(self)
        if name in cls._fields:
            k = cls._fkinds.get(name)
            if k == 'f' and type(value) is int:
                value = float(value)
            elif k == 'c' and type(value) in (int, float):
                value = complex(value)
            self.__dict__[name] = value
            return
        for e in cls._embedded:
            inner = self.__dict__.get(e)
            if inner is not None and inner._has_field(name):
                setattr(inner, name, value)
                return
        raise AttributeError("%s has no field %s" % (cls.__name__, name))

    def _has_field(self, name):
        if name in type(self)._fields:
            return True
        for e in type(self)._embedded:
            inner = self.__dict__.get(e)
            if inner is not None and inner._has_field(name):
                return True
        return False

    def _clone(self):                                   # struct VALUE copy (fields copied shallowly, embedded structs by value)
        cls = type(self)
        vf = cls.__dict__.get("_vf_cache")
        if vf is None:
            vf = tuple(cls._valfields) + tuple(cls._valfields_lazy())
            cls._vf_cache = vf
        o = cls.__new__(cls)
        for k, v in self.__dict__.items():
            o.__dict__[k] = v._clone() if (k in vf and isinstance(v, GoStruct)) else v
        return o

    _valfields = ()       # fields whose type is a struct VALUE (copied with the struct) ...
    _valfields_lazy = staticmethod(lambda: ())   # ... and those whose type lives in another package (struct or interface?)


class NamedBasic(object):
    """a named non-struct type (type X int32): carries its zero value and conversion"""

    def __init__(self, name, zero, conv):
        self.name, self.zero, self.conv = name, zero, conv

    def __call__(self, v):
        return self.conv(v)


class Interface(object):
    def __init__(self, name):
        self.name = name


# ---------------------------------------------------------------- slices
class GoSlice(object):
    __slots__ = ("a", "o", "n", "c", "k")

    def __init__(self, a, o, n, c, k=None):
        self.a, self.o, self.n, self.c, self.k = a, o, n, c, k

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        if type(i) is not int:
            i = int(i) if isinstance(i, int) else _bad_index(i)
        if i < 0 or i >= self.n:
            raise GoPanic("index out of range [%d] with length %d" % (i, self.n))
        return self.a[self.o + i]

    def __setitem__(self, i, v):
        if i < 0 or i >= self.n:
            raise GoPanic("index out of range [%d] with length %d" % (i, self.n))
        k = self.k
        if k == 'f':
            if type(v) is int:
                v = float(v)
        elif k == 'c':
            if type(v) in (int, float):
                v = complex(v)
        self.a[self.o + i] = v

    def __iter__(self):
        a, o = self.a, self.o
        for i in range(self.n):
            yield a[o + i]

    def tolist(self):
        return self.a[self.o:self.o + self.n]


def _bad_index(i):
    raise GoPanic("non-integer slice index %r" % (i,))


_ZERO_OF_KIND = {'f': 0.0, 'i': 0, 'c': 0j, 's': "", 'b': False}


class _G(object):
    """the helper namespace generated code calls (one global object, so Go identifiers can never shadow a helper)"""
    GoStruct, GoSlice, GoError, GoPanic, NamedBasic, Interface = GoStruct, GoSlice, GoError, GoPanic, NamedBasic, Interface
    intdiv_sites = {}

    # ---- zero values / construction
    @staticmethod
    def zv(t):
        """zero value of a named type given as a run-time object"""
        if isinstance(t, type) and issubclass(t, GoStruct):
            return t()
        if isinstance(t, NamedBasic):
            return t.zero
        return None                                     # interface, func, pointer

    @staticmethod
    def make_slice(zero, n, cap=None, kind=None):
        n = int(n)
        cap = n if cap is None else int(cap)
        if n < 0 or cap < n:
            raise GoPanic("makeslice: len out of range")
        return GoSlice([zero() for _ in range(cap)] if callable(zero) else [zero] * cap, 0, n, cap, kind)

    @staticmethod
    def lit(vals, kind=None):
        if kind == 'f':
            vals = [float(v) if type(v) is int else v for v in vals]
        elif kind == 'c':
            vals = [complex(v) if type(v) in (int, float) else v for v in vals]
        return GoSlice(list(vals), 0, len(vals), len(vals), kind)

    @staticmethod
    def lit_indexed(pairs, zero, kind=None):
        """array / slice literal with explicit indices"""
        n = 1 + max(i for i, _ in pairs) if pairs else 0
        a = [zero() if callable(zero) else zero for _ in range(n)]
        for i, v in pairs:
            a[i] = v
        return _G.lit(a, kind)

    @staticmethod
    def len(x):
        return 0 if x is None else len(x)

    @staticmethod
    def cap(x):
        return 0 if x is None else (x.c - 0 if isinstance(x, GoSlice) else len(x))

    @staticmethod
    def slc(s, lo, hi, mx=None):
        if isinstance(s, str):
            return s[(0 if lo is None else lo):(len(s) if hi is None else hi)]
        if s is None:
            if (lo or 0) == 0 and (hi or 0) == 0:
                return None
            raise GoPanic("slice bounds out of range (nil slice)")
        lo = 0 if lo is None else int(lo)
        hi = s.n if hi is None else int(hi)
        cap = s.c if mx is None else int(mx)
        if lo < 0 or hi < lo or hi > s.c or cap > s.c or hi > cap:
            raise GoPanic("slice bounds out of range [%d:%d] with capacity %d" % (lo, hi, s.c))
        return GoSlice(s.a, s.o + lo, hi - lo, cap - lo, s.k)

    @staticmethod
    def copy(dst, src):
        if dst is None or src is None:
            return 0
        n = min(len(dst), len(src))
        if isinstance(src, str):
            for i in range(n):
                dst[i] = ord(src[i])
            return n
        vals = src.a[src.o:src.o + n]                   # snapshot: overlapping copies behave like memmove
        a, o, byval = dst.a, dst.o, (src.k == 'v' or dst.k == 'v')
        for i in range(n):
            v = vals[i]
            a[o + i] = v._clone() if (byval and isinstance(v, GoStruct)) else v
        return n

    @staticmethod
    def kindof(t):
        """element kind of a slice of the named type t (a run-time object of another package)"""
        return 'v' if (isinstance(t, type) and issubclass(t, GoStruct)) else None

    @staticmethod
    def append(s, *vals):
        if s is not None and s.k == 'v':
            vals = tuple((v._clone() if isinstance(v, GoStruct) else v) for v in vals)
        if s is None:
            return GoSlice(list(vals), 0, len(vals), len(vals), None) if vals else None
        need = s.n + len(vals)
        if need <= s.c:                                  # room in the backing array: Go writes in place
            for i, v in enumerate(vals):
                s.a[s.o + s.n + i] = v
            return GoSlice(s.a, s.o, need, s.c, s.k)
        a = s.a[s.o:s.o + s.n] + list(vals)
        cap = max(need, 2 * s.c)
        a += [None] * (cap - need)
        return GoSlice(a, 0, need, cap, s.k)

    @staticmethod
    def spread(s):
        return () if s is None else tuple(s)

    @staticmethod
    def rng(x):
        """`for i := range x`"""
        if x is None:
            return range(0)
        if isinstance(x, int):
            return range(x)
        return range(len(x))

    @staticmethod
    def rng2(x):
        """`for i, v := range x` - struct elements are copied like in Go"""
        if x is None:
            return ()
        if isinstance(x, dict):
            return list(x.items())
        if isinstance(x, str):
            return list(enumerate(x))
        return _G._rng2_slice(x)

    @staticmethod
    def _rng2_slice(x):
        a, o, vals = x.a, x.o, x.k == 'v'
        for i in range(x.n):                            # the range expression is evaluated once, the elements as they are reached
            v = a[o + i]
            yield i, (v._clone() if (vals and isinstance(v, GoStruct)) else v)

    @staticmethod
    def lit_fixed(s, n, zero, kind=None):
        """array literal [n]T{...}: the listed values, padded with zero values"""
        n = int(n)
        if s.n > n:
            raise GoPanic("array index out of bounds in literal")
        pad = [zero() if callable(zero) else zero for _ in range(n - s.n)]
        return GoSlice(s.a[:s.n] + pad, 0, n, n, kind)

    # ---- arithmetic
    @staticmethod
    def div(a, b, site=None):
        ta, tb = type(a), type(b)
        if (ta is int or ta is bool) and (tb is int or tb is bool):
            if site is not None:
                _G.intdiv_sites[site] = _G.intdiv_sites.get(site, 0) + 1
            if b == 0:
                raise GoPanic("integer divide by zero")
            q = abs(a) // abs(b)
            return q if (a >= 0) == (b >= 0) else -q
        if ta is complex or tb is complex:
            return a / b
        try:
            return a / b
        except ZeroDivisionError:
            a = float(a)
            if a != a or a == 0.0:
                return float("nan")
            neg = (a < 0.0) != (math.copysign(1.0, float(b)) < 0.0)
            return float("-inf") if neg else float("inf")

    @staticmethod
    def mod(a, b):
        if b == 0:
            raise GoPanic("integer divide by zero")
        r = abs(a) % abs(b)
        return r if a >= 0 else -r

    @staticmethod
    def andnot(a, b):
        return a & ~b

    # ---- conversions
    @staticmethod
    def f64(x):
        return float(x)

    @staticmethod
    def f32(x):
        return _struct.unpack("f", _struct.pack("f", float(x)))[0]

    @staticmethod
    def _toint(x):
        if type(x) is float:
            if x != x or x in (float("inf"), float("-inf")):
                return -(1 << 63)                        # amd64 behaviour
            return int(x)
        return int(x)

    @staticmethod
    def cint(x):
        return _G._wrap_s(_G._toint(x), 64)

    @staticmethod
    def _wrap_s(v, bits):
        v &= (1 << bits) - 1
        return v - (1 << bits) if v >> (bits - 1) else v

    @staticmethod
    def i8(x): return _G._wrap_s(_G._toint(x), 8)
    @staticmethod
    def i16(x): return _G._wrap_s(_G._toint(x), 16)
    @staticmethod
    def i32(x): return _G._wrap_s(_G._toint(x), 32)
    @staticmethod
    def i64(x): return _G._wrap_s(_G._toint(x), 64)
    @staticmethod
    def u8(x): return _G._toint(x) & 0xFF
    @staticmethod
    def u16(x): return _G._toint(x) & 0xFFFF
    @staticmethod
    def u32(x): return _G._toint(x) & 0xFFFFFFFF
    @staticmethod
    def u64(x): return _G._toint(x) & 0xFFFFFFFFFFFFFFFF

    @staticmethod
    def c128(x):
        return complex(x)

    @staticmethod
    def cstr(x):
        if isinstance(x, int):
            return chr(x)
        if isinstance(x, GoSlice):
            return bytes(x.tolist()).decode("utf-8", "replace")
        return str(x)

    @staticmethod
    def cmplx(re, im):
        return complex(float(re), float(im))

    @staticmethod
    def conv(t, x):
        """conversion to a named type given as a run-time object"""
        if isinstance(t, NamedBasic):
            return t.conv(x)
        return x

    # ---- type tests
    @staticmethod
    def isa(x, t):
        if x is None:
            return False
        if isinstance(t, type):
            return isinstance(x, t)
        if isinstance(t, Interface):
            return True
        if t == "float64": return type(x) is float
        if t == "string": return type(x) is str
        if t == "bool": return type(x) is bool
        if t in ("int", "int32", "int64", "uint32", "uint64", "uint", "int16", "uint16", "int8", "uint8", "byte"): return type(x) is int
        return False

    @staticmethod
    def assert_type(x, t):
        if not _G.isa(x, t):
            raise GoPanic("interface conversion failed: %r is not %r" % (type(x).__name__, t))
        return x

    @staticmethod
    def panic(v):
        raise GoPanic(v)

    # ---- fmt
    @staticmethod
    def sprintf(fmt, *args):
        out, i, k = [], 0, 0
        while i < len(fmt):
            ch = fmt[i]
            if ch != '%':
                out.append(ch); i += 1; continue
            j = i + 1
            while j < len(fmt) and fmt[j] in "+-# 0123456789.":
                j += 1
            verb = fmt[j] if j < len(fmt) else '%'
            spec = fmt[i + 1:j]
            if verb == '%':
                out.append('%'); i = j + 1; continue
            a = args[k] if k < len(args) else "%!" + verb + "(MISSING)"
            k += 1
            if isinstance(a, GoError):
                a = a.msg
            if verb in "vsd" and not spec:
                out.append("true" if a is True else "false" if a is False else str(a))
            elif verb == 'v':
                out.append(str(a))
            elif verb == 'q':
                out.append('"%s"' % a)
            elif verb in "xXoc":
                out.append(("%" + spec + verb) % a)
            elif verb in "eEfFgG":
                out.append(("%" + spec + verb) % float(a))
            else:
                out.append(("%" + spec + ("d" if verb == 'd' else "s")) % a)
            i = j + 1
        return "".join(out)

    @staticmethod
    def errorf(fmt, *args):
        return GoError(_G.sprintf(fmt, *args))


class Chan(object):
    """a channel as far as controller.process needs one: `handler(value)` stands for the goroutine that serves the channel - it
    runs at once on send; what it returns is queued on `reply` (another Chan) for the next receive"""

    def __init__(self, capacity=0, handler=None, reply=None):
        self.queue, self.handler, self.reply, self.closed = [], handler, reply, False

    def send(self, v):
        if self.handler is not None:
            r = self.handler(v)
            (self.reply if self.reply is not None else self).queue.append(r)
        else:
            self.queue.append(v)

    def recv(self):
        if not self.queue:
            raise GoPanic("receive from an empty channel (no goroutines here: nothing would ever send)")
        return self.queue.pop(0)


_G.Chan = Chan
_G.send = staticmethod(lambda ch, v: ch.send(v))
_G.recv = staticmethod(lambda ch: ch.recv())
_G.close = staticmethod(lambda ch: setattr(ch, "closed", True))

G = _G()


# ---------------------------------------------------------------- standard-library shims (only what the path's packages use)
class _Math(object):
    Pi = math.pi
    E = math.e
    Sqrt2 = math.sqrt(2.0)
    Ln2 = math.log(2.0)
    Ln10 = math.log(10.0)
    MaxFloat64 = 1.79769313486231570814527423731704356798070e+308
    SmallestNonzeroFloat64 = 4.9406564584124654417656879286822137236505980e-324
    MaxInt32, MinInt32, MaxInt64, MinInt64 = (1 << 31) - 1, -(1 << 31), (1 << 63) - 1, -(1 << 63)
    MaxInt16, MinInt16, MaxInt8, MinInt8 = (1 << 15) - 1, -(1 << 15), 127, -128
    MaxUint8, MaxUint16, MaxUint32, MaxUint64 = 255, 65535, (1 << 32) - 1, (1 << 64) - 1

    @staticmethod
    def _f(fn, x, over=float("inf")):
        try:
            return fn(float(x))
        except OverflowError:
            return over
        except ValueError:
            return float("nan")

    def Exp(self, x): return self._f(math.exp, x)
    def Exp2(self, x): return self._f(lambda v: 2.0 ** v, x)
    def Log(self, x):
        x = float(x)
        return float("-inf") if x == 0.0 else self._f(math.log, x)
    def Log10(self, x):
        x = float(x)
        return float("-inf") if x == 0.0 else self._f(math.log10, x)
    def Log2(self, x):
        x = float(x)
        return float("-inf") if x == 0.0 else self._f(math.log2, x)
    def Log1p(self, x): return self._f(math.log1p, x)
    def Sin(self, x): return self._f(math.sin, x)
    def Cos(self, x): return self._f(math.cos, x)
    def Tan(self, x): return self._f(math.tan, x)
    def Sincos(self, x): return (self.Sin(x), self.Cos(x))
    def Asin(self, x): return self._f(math.asin, x)
    def Acos(self, x): return self._f(math.acos, x)
    def Atan(self, x): return self._f(math.atan, x)
    def Atan2(self, y, x): return math.atan2(float(y), float(x))
    def Sinh(self, x): return self._f(math.sinh, x)
    def Cosh(self, x): return self._f(math.cosh, x)
    def Tanh(self, x): return self._f(math.tanh, x)
    def Sqrt(self, x): return self._f(math.sqrt, x)
    def Cbrt(self, x):
        x = float(x)
        return math.copysign(abs(x) ** (1.0 / 3.0), x)
    def Abs(self, x): return abs(float(x))
    def Floor(self, x):
        x = float(x)
        return x if (x != x or x in (float("inf"), float("-inf"))) else float(math.floor(x))
    def Ceil(self, x):
        x = float(x)
        return x if (x != x or x in (float("inf"), float("-inf"))) else float(math.ceil(x))
    def Trunc(self, x):
        x = float(x)
        return x if (x != x or x in (float("inf"), float("-inf"))) else float(math.trunc(x))
    def Round(self, x):                                  # half away from zero
        x = float(x)
        if x != x or x in (float("inf"), float("-inf")):
            return x
        return math.copysign(float(math.floor(abs(x) + 0.5)), x) if abs(x) < 4503599627370496.0 else x
    def RoundToEven(self, x): return float(round(float(x)))
    def Mod(self, x, y):
        x, y = float(x), float(y)
        if y == 0.0 or x in (float("inf"), float("-inf")) or x != x or y != y:
            return float("nan")
        return math.fmod(x, y)
    def Modf(self, x):
        x = float(x)
        ip = self.Trunc(x)
        return (ip, x - ip if ip == ip and abs(ip) != float("inf") else 0.0)
    def Pow(self, x, y):
        try:
            return math.pow(float(x), float(y))
        except OverflowError:
            return float("inf")
        except ValueError:
            return float("nan")
    def Hypot(self, x, y): return math.hypot(float(x), float(y))
    def Max(self, x, y):
        x, y = float(x), float(y)
        return float("nan") if (x != x or y != y) else max(x, y)
    def Min(self, x, y):
        x, y = float(x), float(y)
        return float("nan") if (x != x or y != y) else min(x, y)
    def Inf(self, sign): return float("inf") if sign >= 0 else float("-inf")
    def IsInf(self, x, sign):
        x = float(x)
        return (x == float("inf") and sign >= 0) or (x == float("-inf") and sign <= 0)
    def IsNaN(self, x): return float(x) != float(x)
    def NaN(self): return float("nan")
    def Signbit(self, x): return math.copysign(1.0, float(x)) < 0.0
    def Copysign(self, x, y): return math.copysign(float(x), float(y))
    def Float64bits(self, x): return _struct.unpack("<Q", _struct.pack("<d", float(x)))[0]
    def Float64frombits(self, b): return _struct.unpack("<d", _struct.pack("<Q", b & 0xFFFFFFFFFFFFFFFF))[0]
    def Float32bits(self, x): return _struct.unpack("<I", _struct.pack("<f", float(x)))[0]
    def Float32frombits(self, b): return _struct.unpack("<f", _struct.pack("<I", b & 0xFFFFFFFF))[0]


class _Cmplx(object):
    def Exp(self, z): return cmath.exp(complex(z))
    def Abs(self, z): return abs(complex(z))
    def Phase(self, z): return cmath.phase(complex(z))
    def Conj(self, z): return complex(z).conjugate()
    def Sqrt(self, z): return cmath.sqrt(complex(z))
    def Rect(self, r, t): return cmath.rect(float(r), float(t))
    def Polar(self, z): return cmath.polar(complex(z))
    def Pow(self, a, b): return complex(a) ** complex(b)
    def Log(self, z): return cmath.log(complex(z))
    def Inf(self): return complex(float("inf"), float("inf"))
    def NaN(self): return complex(float("nan"), float("nan"))
    def IsNaN(self, z): return z != z
    def IsInf(self, z): return abs(z.real) == float("inf") or abs(z.imag) == float("inf")


class _Bits(object):
    def Len64(self, x): return int(x).bit_length()
    def Len32(self, x): return int(x).bit_length()
    def Len(self, x): return int(x).bit_length()
    def LeadingZeros64(self, x): return 64 - int(x).bit_length()
    def TrailingZeros64(self, x):
        x = int(x)
        return 64 if x == 0 else (x & -x).bit_length() - 1
    def OnesCount64(self, x): return bin(int(x)).count("1")
    def Reverse64(self, x): return int(format(int(x) & (2 ** 64 - 1), "064b")[::-1], 2)


class _RWMutex(GoStruct):
    def Lock(self): pass
    def Unlock(self): pass
    def RLock(self): pass
    def RUnlock(self): pass


class _Sync(object):
    RWMutex = _RWMutex
    Mutex = _RWMutex


class _Fmt(object):
    def Errorf(self, fmt, *a): return G.errorf(fmt, *a)
    def Sprintf(self, fmt, *a): return G.sprintf(fmt, *a)
    def Sprint(self, *a): return "".join(str(x) for x in a)
    def Println(self, *a): print(*a)
    def Printf(self, fmt, *a): print(G.sprintf(fmt, *a), end="")


class _Errors(object):
    def New(self, msg): return GoError(msg)


class _Strconv(object):
    def ParseUint(self, s, base, bits):
        try:
            v = int(s, base if base else 10)
            if v < 0 or v >= (1 << (bits or 64)) or s[:1] in "+-":
                raise ValueError
            return v, None
        except ValueError:
            return 0, GoError('strconv.ParseUint: parsing "%s": invalid syntax' % s)
    def ParseInt(self, s, base, bits):
        try:
            return int(s, base if base else 10), None
        except ValueError:
            return 0, GoError('strconv.ParseInt: parsing "%s": invalid syntax' % s)
    def ParseFloat(self, s, bits):
        try:
            return float(s), None
        except ValueError:
            return 0.0, GoError('strconv.ParseFloat: parsing "%s": invalid syntax' % s)
    def Atoi(self, s): return self.ParseInt(s, 10, 0)
    def Itoa(self, i): return str(i)
    def FormatInt(self, i, base):
        if base == 10:
            return str(i)
        digs, n, out = "0123456789abcdefghijklmnopqrstuvwxyz", abs(i), ""
        while True:
            out = digs[n % base] + out
            n //= base
            if n == 0:
                break
        return ("-" if i < 0 else "") + out
    def FormatUint(self, i, base): return self.FormatInt(i, base)
    def FormatFloat(self, f, fmtc, prec, bits):
        import numpy as np
        fmtc = chr(fmtc) if isinstance(fmtc, int) else fmtc
        if fmtc == 'f' and prec < 0:
            return np.format_float_positional(float(f), unique=True, trim='-')
        if prec < 0:
            return repr(float(f))
        return ("%." + str(prec) + fmtc) % float(f)


class _BytesBuffer(GoStruct):
    """bytes.Buffer: the write side the translated packages use"""

    def _zero(self):
        self.__dict__["data"] = bytearray()

    def WriteByte(self, b):
        self.data.append(int(b) & 0xFF)
        return None

    def Write(self, p):
        vals = p.tolist() if isinstance(p, GoSlice) else list(p)
        self.data.extend(int(v) & 0xFF for v in vals)
        return len(vals), None

    def Bytes(self):
        return GoSlice(list(self.data), 0, len(self.data), len(self.data), "byte")

    def Len(self):
        return len(self.data)


class _BytesReader(object):
    def __init__(self, data):
        self.data, self.pos = bytes(data.tolist()) if data is not None else b"", 0

    def Read(self, buf):
        n = min(len(buf), len(self.data) - self.pos)
        if n <= 0 and len(buf) > 0:
            return 0, GoError("EOF")
        for i in range(n):
            buf[i] = self.data[self.pos + i]
        self.pos += n
        return n, None

    def Len(self):
        return len(self.data) - self.pos

    def Seek(self, offset, whence):
        base = 0 if whence == 0 else (self.pos if whence == 1 else len(self.data))
        p = base + int(offset)
        if p < 0:
            return 0, GoError("bytes.Reader.Seek: negative position")
        self.pos = p
        return p, None


class _Bytes(object):
    Buffer = _BytesBuffer

    def NewReader(self, data):
        return _BytesReader(data)

    def NewBuffer(self, data):
        b = _BytesBuffer()
        if data is not None:
            b.data.extend(data.tolist())
        return b


_PACK = {"int8": "b", "byte": "B", "uint8": "B", "int16": "h", "uint16": "H", "int32": "i", "uint32": "I", "rune": "i", "int64": "q", "uint64": "Q",
         "float32": "f", "f": "d"}


class _Binary(object):
    """encoding/binary for slices of sized numbers (the sample conversions of wave.go); struct headers are not supported"""
    LittleEndian, BigEndian = "<", ">"

    def Write(self, w, order, data):
        if not isinstance(data, GoSlice) or data.k not in _PACK:
            return GoError("binary.Write: unsupported data (the translation handles slices of sized numbers)")
        vals = data.tolist()
        try:
            w.data.extend(_struct.pack(order + "%d%s" % (len(vals), _PACK[data.k]), *vals))
        except _struct.error as e:
            return GoError("binary.Write: %s" % e)
        return None

    def Read(self, r, order, data):
        if not isinstance(data, GoSlice) or data.k not in _PACK:
            return GoError("binary.Read: unsupported data (the translation handles slices of sized numbers)")
        size = _struct.calcsize(order + _PACK[data.k]) * len(data)
        if len(r.data) - r.pos < size:
            return GoError("unexpected EOF")
        vals = _struct.unpack(order + "%d%s" % (len(data), _PACK[data.k]), r.data[r.pos:r.pos + size])
        r.pos += size
        for i, v in enumerate(vals):
            data.a[data.o + i] = v
        return None


class _Io(object):
    SeekStart, SeekCurrent, SeekEnd = 0, 1, 2


STDLIB = {"bytes": _Bytes(), "encoding/binary": _Binary(), "io": _Io(), "math": _Math(), "math/cmplx": _Cmplx(), "math/bits": _Bits(), "sync": _Sync(), "fmt": _Fmt(), "errors": _Errors(),
          "strconv": _Strconv()}
