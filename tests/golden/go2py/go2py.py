"""go2py - a mechanical translator from the subset of Go that go-dsp-guitar's DSP packages use to Python.

TEST INFRASTRUCTURE.  There is no Go toolchain in this image or on the GPU box, so the reference cannot be executed.  This
translator is the next best thing: it reads the reference's own .go files where they lie (never copied into this
repository), translates them statement by statement with no knowledge of what they compute, and the result is executed to
produce golden vectors (make_reference_vectors.py).  No human reading of the DSP code is involved, so a misreading shared by
the C oracle and the hand-written Python restatement (tests/goref.py) cannot hide here.  The translation itself is checked
by running the reference's OWN unit-test vectors (fft, oversampling, resample, random, circular) through the translated
packages before anything else is generated.

Supported: package-level const (iota) / var / type (struct, interface, named basic) / func / methods; statements: := = op=
++ -- var if/else for (all forms, range) switch (tag, tagless, type switch) return break continue; expressions with Go
precedence, composite literals, slices (aliasing views), maps, type assertions, conversions, variadics, multiple results.
Channels exist as far as controller.process needs them: `ch <- v`, `<-ch`, make(chan T), close - a G.Chan runs a handler on send
(the worker that the reference starts as a goroutine) and hands its result to the next receive.
Not supported (absent from the translated code): goroutines (`go`), select, defer, goto, labels, closures, generics.
See goruntime.py for how values are represented."""
import re
from fractions import Fraction

# Untyped constants are exact in Go (arbitrary precision; rounded once, when they meet a typed operand): constant expressions -
# package-level const declarations and constant sub-expressions inside functions alike - are folded here with exact rationals
# and emitted as the correctly rounded float64 literal.  K = (value, is_float): an int for integer constants, a Fraction
# otherwise.  math's constants as in Go's math/const.go.
KCONST = {"math": {
    "Pi": (Fraction("3.14159265358979323846264338327950288419716939937510582097494459"), True),
    "E": (Fraction("2.71828182845904523536028747135266249775724709369995957496696763"), True),
    "Phi": (Fraction("1.61803398874989484820458683436563811772030917980576286213544862"), True),
    "Sqrt2": (Fraction("1.41421356237309504880168872420969807856967187537694807317667974"), True),
    "SqrtE": (Fraction("1.64872127070012814684865078781416357165377610071014801157507931"), True),
    "SqrtPi": (Fraction("1.77245385090551602729816748334114518279754945612238712821380779"), True),
    "SqrtPhi": (Fraction("1.27201964951406896425242246173749149171560804184009624861664038"), True),
    "Ln2": (Fraction("0.693147180559945309417232121458176568075500134360255254120680009"), True),
    "Ln10": (Fraction("2.30258509299404568401799145468436420760110148862877297603332790"), True),
    "MaxInt8": (127, False), "MinInt8": (-128, False), "MaxInt16": (32767, False), "MinInt16": (-32768, False),
    "MaxInt32": (2147483647, False), "MinInt32": (-2147483648, False), "MaxInt64": ((1 << 63) - 1, False), "MinInt64": (-(1 << 63), False),
    "MaxUint8": (255, False), "MaxUint16": (65535, False), "MaxUint32": ((1 << 32) - 1, False), "MaxUint64": ((1 << 64) - 1, False),
}}
KCONST["math"]["Log2E"] = (1 / KCONST["math"]["Ln2"][0], True)
KCONST["math"]["Log10E"] = (1 / KCONST["math"]["Ln10"][0], True)


def kcode(k):
    """python literal of an exact constant"""
    v, is_float = k
    if is_float:
        return repr(float(v))                               # int / int true division: correctly rounded
    return str(int(v))


PY_RESERVED = {"False", "None", "True", "and", "as", "assert", "async", "await", "class", "def", "del", "elif", "except",
               "finally", "from", "global", "import", "in", "is", "lambda", "nonlocal", "not", "or", "pass", "raise", "try",
               "while", "with", "yield", "G", "self", "print", "exec", "eval"}
GO_KEYWORDS = {"break", "case", "chan", "const", "continue", "default", "defer", "else", "fallthrough", "for", "func", "go",
               "goto", "if", "import", "interface", "map", "package", "range", "return", "select", "struct", "switch", "type", "var"}
INT_TYPES = {"int": "G.cint", "int8": "G.i8", "int16": "G.i16", "int32": "G.i32", "int64": "G.i64", "uint": "G.u64", "uint8": "G.u8",
             "uint16": "G.u16", "uint32": "G.u32", "uint64": "G.u64", "uintptr": "G.u64", "byte": "G.u8", "rune": "G.i32"}
BASIC_CONV = dict(INT_TYPES, float64="G.f64", float32="G.f32", complex128="G.c128", complex64="G.c128", string="G.cstr", bool="bool")
OPS = ["<<=", ">>=", "&^=", "...", "&&", "||", "<-", "++", "--", "==", "!=", "<=", ">=", ":=", "+=", "-=", "*=", "/=", "%=", "&=", "|=",
       "^=", "<<", ">>", "&^", "+", "-", "*", "/", "%", "&", "|", "^", "<", ">", "=", "!", "(", ")", "[", "]", "{", "}", ",", ";", ".", ":", "~"]
BINPREC = {"||": 1, "&&": 2, "==": 3, "!=": 3, "<": 3, "<=": 3, ">": 3, ">=": 3, "+": 4, "-": 4, "|": 4, "^": 4,
           "*": 5, "/": 5, "%": 5, "<<": 5, ">>": 5, "&": 5, "&^": 5}


class GoSyntaxError(Exception):
    pass


class Tok(object):
    __slots__ = ("k", "v", "line")

    def __init__(self, k, v, line):
        self.k, self.v, self.line = k, v, line

    def __repr__(self):
        return "%s:%r@%d" % (self.k, self.v, self.line)


def lex(src, fname="?"):
    toks, i, line, n = [], 0, 1, len(src)

    def need_semi():
        if not toks:
            return False
        t = toks[-1]
        if t.k in ("ident", "int", "float", "imag", "char", "string"):
            return True
        if t.k == "kw":
            return t.v in ("break", "continue", "fallthrough", "return")
        return t.k == "op" and t.v in ("++", "--", ")", "]", "}")

    while i < n:
        c = src[i]
        if c == "\n":
            if need_semi():
                toks.append(Tok("op", ";", line))
            line += 1; i += 1; continue
        if c in " \t\r":
            i += 1; continue
        if src.startswith("//", i):
            j = src.find("\n", i)
            i = n if j < 0 else j
            continue
        if src.startswith("/*", i):
            j = src.find("*/", i + 2)
            if j < 0:
                raise GoSyntaxError("%s:%d: unterminated comment" % (fname, line))
            seg = src[i:j + 2]
            if "\n" in seg and need_semi():
                toks.append(Tok("op", ";", line))
            line += seg.count("\n"); i = j + 2; continue
        if c.isalpha() or c == "_":
            j = i + 1
            while j < n and (src[j].isalnum() or src[j] == "_"):
                j += 1
            w = src[i:j]
            toks.append(Tok("kw" if w in GO_KEYWORDS else "ident", w, line)); i = j; continue
        if c.isdigit() or (c == "." and i + 1 < n and src[i + 1].isdigit()):
            m = re.compile(r"0[xX][0-9a-fA-F_]+|0[bB][01_]+|0[oO][0-7_]+|(?:\d[\d_]*)?\.?\d*(?:[eE][+-]?\d+)?").match(src, i)
            j = m.end()
            text = src[i:j]
            if j < n and src[j] == "i":
                toks.append(Tok("imag", text, line)); j += 1
            elif re.match(r"^(0[xXbBoO])", text) or re.match(r"^[\d_]+$", text):
                toks.append(Tok("int", text, line))
            else:
                toks.append(Tok("float", text, line))
            i = j; continue
        if c == '"':
            j = i + 1
            while src[j] != '"':
                j += 2 if src[j] == "\\" else 1
            toks.append(Tok("string", _unquote(src[i + 1:j]), line)); i = j + 1; continue
        if c == "`":
            j = src.find("`", i + 1)
            raw = src[i + 1:j]
            toks.append(Tok("string", raw.replace("\r", ""), line)); line += raw.count("\n"); i = j + 1; continue
        if c == "'":
            j = i + 1
            while src[j] != "'":
                j += 2 if src[j] == "\\" else 1
            toks.append(Tok("char", _unquote(src[i + 1:j]), line)); i = j + 1; continue
        for op in OPS:
            if src.startswith(op, i):
                toks.append(Tok("op", op, line)); i += len(op); break
        else:
            raise GoSyntaxError("%s:%d: unexpected character %r" % (fname, line, c))
    if need_semi():
        toks.append(Tok("op", ";", line))
    toks.append(Tok("eof", "", line))
    return toks


def _unquote(s):
    out, i = [], 0
    simple = {"n": "\n", "t": "\t", "r": "\r", "\\": "\\", '"': '"', "'": "'", "a": "\a", "b": "\b", "f": "\f", "v": "\v", "0": "\0"}
    while i < len(s):
        c = s[i]
        if c != "\\":
            out.append(c); i += 1; continue
        e = s[i + 1]
        if e == "x":
            out.append(chr(int(s[i + 2:i + 4], 16))); i += 4
        elif e == "u":
            out.append(chr(int(s[i + 2:i + 6], 16))); i += 6
        elif e == "U":
            out.append(chr(int(s[i + 2:i + 10], 16))); i += 10
        elif e in "01234567" and i + 3 < len(s) + 1 and re.match(r"[0-7]{3}", s[i + 1:i + 4]):
            out.append(chr(int(s[i + 1:i + 4], 8))); i += 4
        else:
            out.append(simple[e]); i += 2
    return "".join(out)


def safe(name):
    return name + "_" if name in PY_RESERVED else name


class Ex(object):
    """a translated expression: Python code, a rough static kind ('f' float64, 'i' integer, 'c' complex, 's' string, 'b' bool,
    None unknown) and whether it is an untyped integer constant expression (may become a float by context)"""
    def __init__(self, code, kind=None, intconst=False):
        self.code, self.kind, self.intconst = code, kind, intconst
        self.is_index = self.is_assert = self.is_call = False
        self.ityp = None                # a sized integer type (uint64, int32, ...) when statically known: arithmetic wraps
        self.k = None                   # (exact value, is_float) when the expression is a Go constant expression


def sized_int(t):
    """the name of a sized integer type (arithmetic on it wraps), or None"""
    if t and t[0] == "name" and t[1] in INT_TYPES and t[1] not in ("int", "uint"):
        return t[1]
    return None


def kind_of_type(t):
    if t and t[0] == "name":
        if t[1] in ("float64", "float32"):
            return "f"
        if t[1] in INT_TYPES:
            return "i"
        if t[1] in ("complex128", "complex64"):
            return "c"
        if t[1] == "string":
            return "s"
        if t[1] == "bool":
            return "b"
    return None


class PackageTranslator(object):
    def __init__(self, pkgname, files, tolerant=False):
        """files: list of (path, source); tolerant: a top-level declaration that uses an unsupported construct is left out
        (recorded in self.skipped) instead of failing the package - for packages of which only some functions are wanted"""
        self.tolerant, self.skipped = tolerant, []
        self.pkg = pkgname
        self.files = files
        self.types = {}                 # package-level type name -> 'struct' | 'iface' | 'func' | ('basic', underlying type)
        self.toplevel = set()           # every package-level identifier (and import name)
        self.imports = {}               # local name -> import path
        self.defs, self.inits = [], []  # python chunks
        self.methods = []               # (type, method)
        self.tmp = 0
        self.scopes, self.iota, self.nolit, self.stop_at_type_assert = [], None, 0, False
        self.kconsts = {}               # package-level constants with an exact value: name -> (value, is_float)
        KCONST["github.com/andrepxx/go-dsp-guitar/" + pkgname] = self.kconsts
        KCONST[pkgname] = self.kconsts
        for path, src in files:
            for m in re.finditer(r"^type\s+(\w+)\s+(\S+)", src, re.M):
                self._note_type(m.group(1), m.group(2))
            for blk in re.finditer(r"^type\s*\((.*?)^\)", src, re.M | re.S):
                for m in re.finditer(r"^\s+(\w+)\s+(\S+)", blk.group(1), re.M):
                    self._note_type(m.group(1), m.group(2))
            for m in re.finditer(r"^func\s+(\w+)\s*\(", src, re.M):
                self.toplevel.add(m.group(1))
            for m in re.finditer(r"^(?:var|const)\s+(\w+)", src, re.M):
                self.toplevel.add(m.group(1))
            for blk in re.finditer(r"^(?:var|const)\s*\((.*?)^\)", src, re.M | re.S):
                for m in re.finditer(r"^\s+(\w+(?:\s*,\s*\w+)*)", blk.group(1), re.M):
                    for nm in m.group(1).split(","):
                        self.toplevel.add(nm.strip())

    def _note_type(self, name, what):
        self.toplevel.add(name)
        if what.startswith("struct"):
            self.types[name] = "struct"
        elif what.startswith("interface"):
            self.types[name] = "iface"
        elif what.startswith("func"):
            self.types[name] = "func"
        else:
            self.types[name] = ("basic", what)

    # ------------------------------------------------------------------ driver
    def translate(self):
        for path, src in self.files:
            self.fname = path
            self.toks = lex(src, path)
            self.p = 0
            self.parse_file()
        return self

    def module_source(self):
        return "\n\n".join(self.defs), list(self.inits)

    # ------------------------------------------------------------------ token helpers
    def peek(self, o=0):
        return self.toks[min(self.p + o, len(self.toks) - 1)]

    def at(self, v, k="op"):
        t = self.toks[self.p]
        return t.k == k and t.v == v

    def at_kw(self, v):
        return self.at(v, "kw")

    def next(self):
        t = self.toks[self.p]
        self.p += 1
        return t

    def expect(self, v, k="op"):
        t = self.next()
        if t.k != k or t.v != v:
            self.err("expected %r, got %r" % (v, t.v), t)
        return t

    def accept(self, v, k="op"):
        if self.at(v, k):
            self.p += 1
            return True
        return False

    def ident(self):
        t = self.next()
        if t.k != "ident":
            self.err("expected identifier, got %r" % (t.v,), t)
        return t.v

    def err(self, msg, t=None):
        t = t or self.peek()
        raise GoSyntaxError("%s:%d: %s" % (self.fname, t.line, msg))

    def skip_semis(self):
        while self.accept(";"):
            pass

    def newtmp(self, base="_t"):
        self.tmp += 1
        return "%s%d" % (base, self.tmp)

    # ------------------------------------------------------------------ file level
    def parse_file(self):
        self.skip_semis()
        self.expect("package", "kw"); self.ident(); self.skip_semis()
        while not self.peek().k == "eof":
            t = self.peek()
            if t.k != "kw":
                self.err("unexpected token %r at top level" % (t.v,))
            start, ndefs, ninits = self.p, len(self.defs), len(self.inits)
            try:
                if t.v == "import":
                    self.parse_import()
                elif t.v == "const":
                    self.parse_const_decl(toplevel=True)
                elif t.v == "var":
                    self.parse_var_decl(toplevel=True)
                elif t.v == "type":
                    self.parse_type_decl()
                elif t.v == "func":
                    self.parse_func()
                else:
                    self.err("unsupported top-level declaration %r" % (t.v,))
            except GoSyntaxError as e:
                if not self.tolerant:
                    raise
                del self.defs[ndefs:]
                del self.inits[ninits:]
                self.scopes, self.nolit, self.iota = [], 0, None
                self.p = start
                depth = 0
                while self.peek().k != "eof":                # resynchronise: the `;` that ends this declaration at depth 0
                    tk = self.next()
                    if tk.k == "op" and tk.v in ("{", "(", "["):
                        depth += 1
                    elif tk.k == "op" and tk.v in ("}", ")", "]"):
                        depth -= 1
                    elif tk.k == "op" and tk.v == ";" and depth == 0 and self.peek().k in ("kw", "eof") \
                            and self.peek().v in ("func", "var", "const", "type", "import", ""):
                        break
                what = " ".join(x.v for x in self.toks[start:start + 6])
                self.skipped.append((what, str(e)))
            self.skip_semis()

    def parse_import(self):
        self.expect("import", "kw")

        def one():
            alias = None
            if self.peek().k == "ident":
                alias = self.ident()
            elif self.accept("."):
                self.err("dot imports are not supported")
            path = self.next().v
            name = alias or path.split("/")[-1]
            self.imports[name] = path
            self.toplevel.add(name)
        if self.accept("("):
            self.skip_semis()
            while not self.at(")"):
                one(); self.skip_semis()
            self.expect(")")
        else:
            one()

    # ------------------------------------------------------------------ types
    def parse_type(self):
        t = self.peek()
        if t.k == "ident":
            name = self.ident()
            if self.at(".") and name in self.imports:
                self.next()
                return ("qual", name, self.ident())
            return ("name", name)
        if t.k == "op" and t.v == "*":
            self.next()
            return ("ptr", self.parse_type())
        if t.k == "op" and t.v == "(":
            self.next(); ty = self.parse_type(); self.expect(")")
            return ty
        if t.k == "op" and t.v == "[":
            self.next()
            if self.accept("]"):
                return ("slice", self.parse_type())
            if self.accept("..."):
                self.expect("]")
                return ("array", None, self.parse_type())
            n = self.parse_expr()
            self.expect("]")
            return ("array", n.code, self.parse_type())
        if t.k == "kw" and t.v == "map":
            self.next(); self.expect("["); k = self.parse_type(); self.expect("]")
            return ("map", k, self.parse_type())
        if t.k == "kw" and t.v == "struct":
            return ("struct", self.parse_struct_body())
        if t.k == "kw" and t.v == "interface":
            self.next(); self.expect("{")
            depth = 1
            while depth:
                x = self.next()
                if x.k == "op" and x.v == "{":
                    depth += 1
                elif x.k == "op" and x.v == "}":
                    depth -= 1
            return ("iface",)
        if t.k == "kw" and t.v == "func":
            self.next(); self.parse_signature()
            return ("func",)
        if t.k == "kw" and t.v == "chan":
            self.next()
            self.accept("<-")
            self.parse_type()
            return ("chan",)
        if t.k == "op" and t.v == "<-":
            self.next(); self.expect("chan", "kw"); self.parse_type()
            return ("chan",)
        self.err("expected a type, got %r" % (t.v,))

    def parse_struct_body(self):
        self.expect("struct", "kw"); self.expect("{"); self.skip_semis()
        fields = []                                         # (name, type, embedded)
        while not self.at("}"):
            if self.peek().k == "op" and self.peek().v == "*":
                self.next()
                ty = self.parse_type()
                fields.append((ty[-1], ("ptr", ty), True))
            else:
                names = [self.ident()]
                if self.at(";") or self.at("}") or self.at(".") or self.peek().k == "string":
                    if self.accept("."):                    # embedded qualified type
                        ty = ("qual", names[0], self.ident())
                        fields.append((ty[2], ty, True))
                    else:
                        fields.append((names[0], ("name", names[0]), True))
                else:
                    while self.accept(","):
                        names.append(self.ident())
                    ty = self.parse_type()
                    for nm in names:
                        fields.append((nm, ty, False))
            if self.peek().k == "string":
                self.next()                                 # field tag
            self.skip_semis()
        self.expect("}")
        return fields

    def kind_code(self, t):
        """python expression for the element kind of a slice of t: 'f' / 'c' (literal coercion), 'v' (struct VALUES: copied by
        range / copy / append), None"""
        k = kind_of_type(t)
        if t and t[0] == "name" and (t[1] == "float32" or (t[1] in INT_TYPES and t[1] not in ("int", "uint"))):
            return repr("byte" if t[1] == "uint8" else t[1])  # sized elements: what encoding/binary needs to know
        if k in ("f", "c"):
            return repr(k)
        if t and t[0] == "name" and self.types.get(t[1]) == "struct":
            return "'v'"
        if t and t[0] == "qual":
            return "G.kindof(%s.%s)" % (safe(t[1]), safe(t[2]))
        return "None"

    def zero_code(self, t, factory=False):
        """python expression for the zero value of type t; with factory=True a callable (or constant) for G.make_slice"""
        k = t[0]
        z = None
        if k == "name":
            n = t[1]
            if n in ("float64", "float32"):
                z = "0.0"
            elif n in INT_TYPES:
                z = "0"
            elif n in ("complex128", "complex64"):
                z = "0j"
            elif n == "string":
                z = "''"
            elif n == "bool":
                z = "False"
            elif n == "error":
                z = "None"
            elif n in self.types:
                what = self.types[n]
                if what == "struct":
                    return ("(lambda: %s())" % safe(n)) if factory else "%s()" % safe(n)
                if isinstance(what, tuple):
                    return self.zero_code(("name", what[1]) if re.match(r"^\w+$", what[1]) else ("slice", None), factory)
                z = "None"
            else:
                z = "None"
        elif k == "qual":
            c = "G.zv(%s.%s)" % (safe(t[1]), safe(t[2]))
            return ("(lambda: %s)" % c) if factory else c
        elif k == "array":
            c = "G.make_slice(%s, %s, None, %s)" % (self.zero_code(t[2], True), t[1], self.kind_code(t[2]))
            return ("(lambda: %s)" % c) if factory else c
        elif k == "struct":
            self.err("anonymous struct types are not supported")
        else:
            z = "None"
        return z

    def parse_type_decl(self):
        self.expect("type", "kw")

        def one():
            name = self.ident()
            self.accept("=")
            if self.at_kw("struct"):
                self.emit_struct(name, self.parse_struct_body())
            else:
                ty = self.parse_type()
                if ty[0] in ("iface", "func"):
                    self.defs.append("%s = G.Interface(%r)" % (safe(name), name))
                elif ty[0] == "name" and ty[1] in BASIC_CONV:
                    self.defs.append("%s = G.NamedBasic(%r, %s, %s)" % (safe(name), name, self.zero_code(ty), BASIC_CONV[ty[1]]))
                else:
                    self.defs.append("%s = G.NamedBasic(%r, None, (lambda v: v))" % (safe(name), name))
        if self.accept("("):
            self.skip_semis()
            while not self.at(")"):
                one(); self.skip_semis()
            self.expect(")")
        else:
            one()

    def emit_struct(self, name, fields):
        names = [safe(f[0]) for f in fields]
        emb = [safe(f[0]) for f in fields if f[2]]
        fk = dict((safe(f[0]), kind_of_type(f[1])) for f in fields if kind_of_type(f[1]) in ("f", "c"))
        valf = [safe(f[0]) for f in fields if (f[1][0] == "name" and self.types.get(f[1][1]) == "struct")]
        # a field of a type from another package is a struct VALUE only if that type is a struct: decided when first needed
        lazy = ", ".join("(%r, %s.%s)" % (safe(f[0]), safe(f[1][1]), safe(f[1][2])) for f in fields if f[1][0] == "qual")
        lines = ["class %s(G.GoStruct):" % safe(name),
                 "    _fields = %r" % (tuple(names),), "    _embedded = %r" % (tuple(emb),), "    _fkinds = %r" % (fk,),
                 "    _valfields = %r" % (tuple(valf),),
                 "    _valfields_lazy = staticmethod(lambda: tuple(n for n, t in (%s) if G.kindof(t) == 'v'))" % (lazy + "," if lazy else ""),
                 "    def _zero(self):", "        d = self.__dict__"]
        for f in fields:
            lines.append("        d[%r] = %s" % (safe(f[0]), self.zero_code(f[1])))
        if not fields:
            lines.append("        pass")
        self.defs.append("\n".join(lines))

    # ------------------------------------------------------------------ const / var
    def parse_const_decl(self, toplevel):
        self.expect("const", "kw")
        out = []

        def spec(iota, prev):
            names = [self.ident()]
            while self.accept(","):
                names.append(self.ident())
            ty = None
            if not self.at("=") and not self.at(";") and not self.at(")"):
                ty = self.parse_type()
            if self.accept("="):
                start = self.p
                self.iota = iota
                vals = [self.parse_expr()]
                while self.accept(","):
                    vals.append(self.parse_expr())
                prev = (ty, start, self.p)
            else:
                if prev is None:
                    self.err("constant without a value")
                ty, s0, e0 = prev
                save = self.p
                self.p = s0
                self.iota = iota
                vals = [self.parse_expr()]
                while self.accept(","):
                    vals.append(self.parse_expr())
                self.p = save
            self.iota = None
            for nm, v in zip(names, vals):
                code = v.code
                kind = kind_of_type(ty) if ty else (v.kind if not v.intconst else "i")
                if toplevel and v.k is not None and nm != "_":
                    if ty is None:
                        self.kconsts[nm] = v.k
                    elif kind == "f":
                        self.kconsts[nm] = (Fraction(float(v.k[0])), True)   # typed float constant: rounded at its declaration
                    elif kind == "i" and not v.k[1]:
                        self.kconsts[nm] = v.k
                if ty and kind == "f":
                    code = "G.f64(%s)" % code
                elif ty and ty[0] == "name" and ty[1] in INT_TYPES and v.kind == "f":
                    code = "%s(%s)" % (INT_TYPES[ty[1]], code)
                if nm == "_":
                    continue
                if toplevel:
                    self.inits.append("%s = %s" % (safe(nm), code))
                else:
                    py = self.declare(nm, kind)
                    out.append("%s = %s" % (py, code))
            return prev
        if self.accept("("):
            self.skip_semis()
            iota, prev = 0, None
            while not self.at(")"):
                prev = spec(iota, prev); iota += 1; self.skip_semis()
            self.expect(")")
        else:
            spec(0, None)
        return out

    def parse_var_decl(self, toplevel):
        self.expect("var", "kw")
        out = []

        def spec():
            names = [self.ident()]
            while self.accept(","):
                names.append(self.ident())
            ty = None
            if not self.at("="):
                ty = self.parse_type()
            vals = None
            if self.accept("="):
                vals = [self.parse_expr()]
                while self.accept(","):
                    vals.append(self.parse_expr())
            kind = kind_of_type(ty) if ty else None
            if vals is None:
                codes = [self.zero_code(ty) for _ in names]
                kinds = [kind] * len(names)
            elif len(vals) == len(names):
                codes = [self.coerce(v, kind).code for v in vals]
                kinds = [kind or ("i" if v.intconst else v.kind) for v in vals]
            else:                                           # var a, b = f()
                codes, kinds = None, [kind] * len(names)
            if toplevel:
                pys = [safe(n) for n in names]
            else:
                pys = [self.declare(n, k, sized_int(ty) if ty else None) for n, k in zip(names, kinds)]
            if codes is None:
                line = "%s = %s" % (", ".join(pys), vals[0].code)
            else:
                line = "%s = %s" % (", ".join(pys), ", ".join(codes))
            (self.inits if toplevel else out).append(line)
        if self.accept("("):
            self.skip_semis()
            while not self.at(")"):
                spec(); self.skip_semis()
            self.expect(")")
        else:
            spec()
        return out

    def coerce(self, e, kind):
        """expression e used where a value of static kind `kind` is expected"""
        if kind == "f" and e.kind != "f" and (e.intconst or e.kind == "i"):
            return Ex("G.f64(%s)" % e.code, "f")
        if kind == "c" and e.kind != "c" and (e.intconst or e.kind in ("i", "f")):
            return Ex("complex(%s)" % e.code, "c")
        return e

    # ------------------------------------------------------------------ functions
    def parse_signature(self):
        """returns (params [(name|None, type, variadic)], results [(name|None, type)])"""
        params = self.parse_param_list()
        results = []
        if self.at("("):
            results = [(n, t) for (n, t, _) in self.parse_param_list()]
        elif not self.at("{") and not self.at(";") and not self.at(")") and not self.at(",") and not self.at("}") \
                and not self.peek().k == "string" and not self.peek().k == "eof" and not self.at("=") and not self.at("]"):
            results = [(None, self.parse_type())]
        return params, results

    def parse_param_list(self):
        self.expect("(")
        entries = []                                        # raw: list of (maybe_name_tokens, type)
        groups = []
        while not self.at(")"):
            variadic = False
            start = self.p
            # try "name Type" / "name, name Type" / "Type"
            if self.peek().k == "ident" and not (self.peek(1).k == "op" and self.peek(1).v in (",", ")", ".")):
                name = self.ident()
                if self.accept("..."):
                    variadic = True
                ty = self.parse_type()
                groups.append(([name], ty, variadic))
            elif self.peek().k == "ident" and self.peek(1).k == "op" and self.peek(1).v == ",":
                # either a list of names followed by a type, or a list of unnamed types
                names = [self.ident()]
                while self.accept(","):
                    if self.peek().k == "ident" and self.peek(1).k == "op" and self.peek(1).v in (",", ")"):
                        names.append(self.ident())
                    elif self.peek().k == "ident" and not (self.peek(1).k == "op" and self.peek(1).v == "."):
                        names.append(self.ident())
                        if self.accept("..."):
                            variadic = True
                        ty = self.parse_type()
                        groups.append((names, ty, variadic))
                        names = None
                        break
                    else:
                        self.p = start
                        names = None
                        ty = self.parse_type()
                        groups.append(([None], ty, False))
                        break
                if names is not None:                       # all identifiers: unnamed parameter types
                    for nm in names:
                        groups.append(([None], ("name", nm), False))
            else:
                if self.accept("..."):
                    variadic = True
                ty = self.parse_type()
                groups.append(([None], ty, variadic))
            if not self.accept(","):
                break
        self.expect(")")
        for names, ty, variadic in groups:
            for nm in names:
                entries.append((nm, ty, variadic))
        return entries

    def parse_func(self):
        self.expect("func", "kw")
        recv = None
        if self.at("("):
            rl = self.parse_param_list()
            rname, rtype, _ = rl[0]
            while rtype[0] == "ptr":
                rtype = rtype[1]
            recv = (rname or "_recv", rtype[1])
        name = self.ident()
        params, results = self.parse_signature()
        if not self.at("{"):
            return                                          # declaration without a body
        self.fn_results = results
        self.fn_globals = set()
        self.scopes = [{}]
        self.fn_used = set()
        self.loop_stack = []
        pyparams, prologue = [], []
        if recv:
            pyparams.append(self.declare(recv[0], None))
        for nm, ty, variadic in params:
            py = self.declare(nm if nm and nm != "_" else self.newtmp("_arg"), kind_of_type(ty) if not variadic else None,
                              sized_int(ty) if not variadic else None)
            if variadic:
                pyparams.append("*" + py)
                prologue.append("%s = G.lit(list(%s), %s)" % (py, py, self.kind_code(ty)))
            else:
                pyparams.append(py)
                if kind_of_type(ty) == "f":
                    prologue.append("if type(%s) is int: %s = float(%s)" % (py, py, py))
        named_results = [r for r in results if r[0]]
        self.named_results = None
        if named_results:
            self.named_results = []
            for nm, ty in results:
                py = self.declare(nm, kind_of_type(ty))
                prologue.append("%s = %s" % (py, self.zero_code(ty)))
                self.named_results.append(py)
        body = self.parse_block(new_scope=False)
        pyname = ("%s__%s" % (safe(recv[1]), safe(name))) if recv else safe(name)
        lines = ["def %s(%s):" % (pyname, ", ".join(pyparams))]
        if self.fn_globals:
            lines.append("    global %s" % ", ".join(sorted(self.fn_globals)))
        for l in prologue + body:
            lines.append("    " + l)
        if len(lines) == 1:
            lines.append("    pass")
        if self.named_results is not None:
            lines.append("    return %s" % ", ".join(self.named_results))
        src = "\n".join(lines)
        self.defs.append(src)
        if recv:
            self.defs.append("%s.%s = %s" % (safe(recv[1]), safe(name), pyname))
        self.scopes = []

    # ------------------------------------------------------------------ scopes
    def declare(self, name, kind, ityp=None):
        if name == "_":
            return "_"
        cur = self.scopes[-1]
        if name in cur:
            cur[name] = (cur[name][0], kind if kind else cur[name][1], ityp or cur[name][2])
            return cur[name][0]
        base = safe(name)
        py = base
        k = 0
        while py in self.fn_used or py in self.toplevel or py in self.imports:
            k += 1
            py = "%s_%d" % (base, k)
        self.fn_used.add(py)
        cur[name] = (py, kind, ityp)
        return py

    def lookup(self, name):
        for sc in reversed(self.scopes):
            if name in sc:
                return sc[name]
        return None

    # ------------------------------------------------------------------ statements
    def parse_block(self, new_scope=True):
        self.expect("{")
        if new_scope:
            self.scopes.append({})
        out = []
        self.skip_semis()
        while not self.at("}"):
            out.extend(self.parse_stmt())
            self.skip_semis()
        self.expect("}")
        if new_scope:
            self.scopes.pop()
        return out or ["pass"]

    def indent(self, lines):
        return ["    " + l for l in lines]

    def parse_stmt(self):
        t = self.peek()
        if t.k == "kw":
            v = t.v
            if v == "var":
                return self.parse_var_decl(toplevel=False)
            if v == "const":
                return self.parse_const_decl(toplevel=False)
            if v == "return":
                return self.parse_return()
            if v == "if":
                return self.parse_if()
            if v == "for":
                return self.parse_for()
            if v == "switch":
                return self.parse_switch()
            if v == "break":
                self.next()
                if self.peek().k == "ident":
                    self.err("labeled break is not supported")
                if self.loop_stack and self.loop_stack[-1][0] == "switch":
                    q = self.p
                    while self.toks[q].k == "op" and self.toks[q].v == ";":
                        q += 1
                    nx = self.toks[q]
                    if (nx.k == "kw" and nx.v in ("case", "default")) or (nx.k == "op" and nx.v == "}"):
                        return []                           # a break that ends its clause: nothing to translate
                    self.err("break in the middle of a switch clause is not supported")
                if not self.loop_stack:
                    self.err("break outside a loop")
                return ["break"]
            if v == "continue":
                self.next()
                if self.peek().k == "ident":
                    self.err("labeled continue is not supported")
                loops = [l for l in self.loop_stack if l[0] == "for"]
                if not loops:
                    self.err("continue outside a loop")
                return list(loops[-1][1]) + ["continue"]
            if v == "type":
                self.err("local type declarations are not supported")
            if v == "func":
                self.err("function literals are not supported")
            self.err("unsupported statement %r" % (v,))
        if t.k == "op" and t.v == "{":
            return self.parse_block()
        if t.k == "op" and t.v == ";":
            self.next()
            return []
        return self.parse_simple_stmt()

    def parse_return(self):
        self.expect("return", "kw")
        if self.at(";") or self.at("}"):
            if self.named_results is not None:
                return ["return %s" % ", ".join(self.named_results)]
            return ["return"]
        vals = [self.parse_expr()]
        while self.accept(","):
            vals.append(self.parse_expr())
        if len(vals) == len(self.fn_results):
            vals = [self.coerce(v, kind_of_type(r[1])) for v, r in zip(vals, self.fn_results)]
        return ["return %s" % ", ".join(v.code for v in vals)]

    def parse_simple_stmt(self):
        """expression statement, inc/dec, assignment, short variable declaration; returns python lines"""
        lhs = [self.parse_expr()]
        while self.accept(","):
            lhs.append(self.parse_expr())
        t = self.peek()
        if t.k == "op" and t.v == "<-" and len(lhs) == 1:        # ch <- v
            self.next()
            v = self.parse_expr()
            return ["G.send(%s, %s)" % (lhs[0].code, v.code)]
        if t.k == "op" and (t.v in ("++", "--", "=") or (t.v.endswith("=") and t.v not in ("==", "!=", "<=", ">=", ":="))):
            for l in lhs:
                if getattr(l, "is_global", False) and l.code != "_":
                    self.fn_globals.add(l.code)
        if t.k == "op" and t.v in ("++", "--"):
            self.next()
            return ["%s %s= 1" % (lhs[0].code, "+" if t.v == "++" else "-")]
        if t.k == "op" and t.v == ":=":
            self.next()
            if self.at_kw("range"):
                self.err("range outside a for clause")
            rhs = [self.parse_expr()]
            while self.accept(","):
                rhs.append(self.parse_expr())
            # the left side was parsed as expressions: recover the identifiers
            names = [self._lhs_name(e) for e in lhs]
            return self.emit_define(names, rhs)
        if t.k == "op" and t.v == "=":
            self.next()
            rhs = [self.parse_expr()]
            while self.accept(","):
                rhs.append(self.parse_expr())
            if len(lhs) == 2 and len(rhs) == 1:
                rhs = self.comma_ok(rhs)
            if len(lhs) == len(rhs):
                rhs = [self.coerce(r, l.kind) for l, r in zip(lhs, rhs)]
            return ["%s = %s" % (", ".join(("_" if l.code == "_" else l.code) for l in lhs), ", ".join(r.code for r in rhs))]
        if t.k == "op" and t.v.endswith("=") and t.v not in ("==", "!=", "<=", ">="):
            self.next()
            op = t.v[:-1]
            r = self.parse_expr()
            l = lhs[0]
            if op == "/":
                return ["%s = G.div(%s, %s, %r)" % (l.code, l.code, r.code, self.site(t))]
            if op == "%":
                return ["%s = G.mod(%s, %s)" % (l.code, l.code, r.code)]
            if op == "&^":
                return ["%s = G.andnot(%s, %s)" % (l.code, l.code, r.code)]
            return ["%s %s= %s" % (l.code, op, r.code)]
        if len(lhs) != 1:
            self.err("unexpected expression list")
        return [lhs[0].code]

    def site(self, t):
        return "%s:%d" % (self.fname.split("/")[-1], t.line)

    def _lhs_name(self, e):
        c = e.code
        if not re.match(r"^\w+$", c):
            self.err("left side of := is not an identifier: %s" % c)
        return getattr(e, "goname", c)

    def comma_ok(self, rhs):
        r = rhs[0]
        if r.is_assert:
            return [Ex("(lambda _v: (_v, G.isa(_v, %s)))(%s)" % (r.assert_type, r.assert_operand))]
        if r.is_index:
            return [Ex("(lambda _m, _k: (_m.get(_k, %s), (_k in _m)) if _m is not None else (%s, False))(%s, %s)"
                       % (r.index_zero, r.index_zero, r.index_base, r.index_key))]
        return rhs

    def emit_define(self, names, rhs):
        if len(names) == 2 and len(rhs) == 1:
            rhs = self.comma_ok(rhs)
        if len(rhs) == len(names):
            kinds = [("i" if r.intconst else r.kind) for r in rhs]
            ityps = [r.ityp for r in rhs]
        else:
            kinds = [None] * len(names)
            ityps = [None] * len(names)
        pys = [self.declare(n, k, it) for n, k, it in zip(names, kinds, ityps)]
        return ["%s = %s" % (", ".join(pys), ", ".join(r.code for r in rhs))]

    def parse_if(self):
        self.expect("if", "kw")
        self.scopes.append({})
        pre = []
        self.nolit = self.nolit + 1
        save = self.p
        first = self.parse_simple_stmt_or_expr()
        if self.accept(";"):
            pre = first[0]
            cond = self.parse_expr()
        else:
            cond = first[1]
            if cond is None:
                self.err("if without a condition")
        self.nolit -= 1
        body = self.parse_block()
        out = pre + ["if %s:" % cond.code] + self.indent(body)
        if self.accept("else", "kw"):
            if self.at_kw("if"):
                els = self.parse_if()
            else:
                els = self.parse_block()
            out += ["else:"] + self.indent(els)
        self.scopes.pop()
        return out

    def parse_simple_stmt_or_expr(self):
        """for if / switch headers: returns (statement lines, expression or None)"""
        start = self.p
        e = self.parse_expr()
        if self.at(";") or self.at("{"):
            if self.at("{"):
                return (None, e)
            # `x; cond` with x an expression statement
            return ([e.code], None)
        self.p = start
        return (self.parse_simple_stmt(), None)

    def parse_for(self):
        self.expect("for", "kw")
        self.scopes.append({})
        self.nolit = self.nolit + 1
        out = []
        if self.at("{"):
            self.nolit -= 1
            self.loop_stack.append(("for", []))
            body = self.parse_block()
            self.loop_stack.pop()
            self.scopes.pop()
            return ["while True:"] + self.indent(body)
        if self.at_kw("range"):
            self.next()
            x = self.parse_expr()
            self.nolit -= 1
            self.loop_stack.append(("for", []))
            body = self.parse_block()
            self.loop_stack.pop()
            self.scopes.pop()
            return ["for _ in G.rng(%s):" % x.code] + self.indent(body)
        # range clause with variables?
        save = self.p
        is_range = False
        depth = 0
        q = self.p
        while True:
            tk = self.toks[q]
            if tk.k == "eof":
                break
            if tk.k == "op" and tk.v in ("(", "["):
                depth += 1
            elif tk.k == "op" and tk.v in (")", "]"):
                depth -= 1
            elif tk.k == "op" and tk.v in ("{", ";") and depth == 0:
                break
            elif tk.k == "kw" and tk.v == "range" and depth == 0:
                is_range = True
                break
            q += 1
        if is_range:
            lhs = [self.parse_expr()]
            while self.accept(","):
                lhs.append(self.parse_expr())
            define = self.next().v == ":="
            self.expect("range", "kw")
            x = self.parse_expr()
            self.nolit -= 1
            if define:
                pys = [self.declare(self._lhs_name(e), None) for e in lhs]
            else:
                pys = [e.code for e in lhs]
            self.loop_stack.append(("for", []))
            body = self.parse_block()
            self.loop_stack.pop()
            self.scopes.pop()
            if len(pys) == 1:
                return ["for %s in G.rng(%s):" % (pys[0], x.code)] + self.indent(body)
            return ["for %s, %s in G.rng2(%s):" % (pys[0], pys[1], x.code)] + self.indent(body)
        # condition-only or three-clause
        init, cond, post = [], None, []
        if self.at(";"):
            first = ([], None)
        else:
            first = self.parse_simple_stmt_or_expr()
        if self.at("{"):                                     # for cond {
            cond = first[1]
            if cond is None:
                self.err("bad for clause")
        else:
            self.expect(";")
            init = first[0] if first[0] is not None else [first[1].code]
            if not self.at(";"):
                cond = self.parse_expr()
            self.expect(";")
            if not self.at("{"):
                post = self.parse_simple_stmt()
        self.nolit -= 1
        self.loop_stack.append(("for", post))
        body = self.parse_block()
        self.loop_stack.pop()
        self.scopes.pop()
        return init + ["while %s:" % (cond.code if cond else "True")] + self.indent(body + post)

    def parse_switch(self):
        self.expect("switch", "kw")
        self.scopes.append({})
        self.nolit = self.nolit + 1
        pre, tag, type_switch, bind = [], None, None, None
        if not self.at("{"):
            # forms: `switch x {`, `switch init; x {`, `switch x := y.(type) {`, `switch y.(type) {`
            start = self.p
            if self.peek().k == "ident" and self.peek(1).k == "op" and self.peek(1).v == ":=" and self._is_type_switch(self.p + 2):
                bind = self.ident(); self.next()
            if self._is_type_switch(self.p):
                e = self.parse_primary_no_assert()
                self.expect("."); self.expect("("); self.expect("type", "kw"); self.expect(")")
                type_switch = e
            else:
                first = self.parse_simple_stmt_or_expr()
                if self.accept(";"):
                    pre = first[0] if first[0] is not None else [first[1].code]
                    if not self.at("{"):
                        tag = self.parse_expr()
                else:
                    tag = first[1]
        self.nolit -= 1
        self.expect("{")
        out = list(pre)
        tv = None
        if tag is not None:
            tv = self.newtmp("_sw")
            out.append("%s = %s" % (tv, tag.code))
        if type_switch is not None:
            tv = self.newtmp("_ts")
            out.append("%s = %s" % (tv, type_switch.code))
        self.skip_semis()
        first_clause = True
        default_body = None
        self.loop_stack.append(("switch", None))
        while not self.at("}"):
            if self.accept("default", "kw"):
                self.expect(":")
                conds = None
            else:
                self.expect("case", "kw")
                conds = []
                if type_switch is not None:
                    tys = [self.parse_type()]
                    while self.accept(","):
                        tys.append(self.parse_type())
                    for ty in tys:
                        if ty == ("name", "nil"):
                            conds.append("(%s is None)" % tv)
                        else:
                            conds.append("G.isa(%s, %s)" % (tv, self.type_object_code(ty)))
                else:
                    es = [self.parse_expr()]
                    while self.accept(","):
                        es.append(self.parse_expr())
                    for e in es:
                        conds.append("(%s == %s)" % (tv, e.code) if tv else e.code)
                self.expect(":")
            self.scopes.append({})
            body = []
            if bind and type_switch is not None:
                body.append("%s = %s" % (self.declare(bind, None), tv))
            self.skip_semis()
            while not (self.at_kw("case") or self.at_kw("default") or self.at("}")):
                if self.at_kw("fallthrough"):
                    self.err("fallthrough is not supported")
                body.extend(self.parse_stmt())
                self.skip_semis()
            self.scopes.pop()
            body = body or ["pass"]
            if conds is None:
                default_body = body
            else:
                out.append("%s %s:" % ("if" if first_clause else "elif", " or ".join(conds)))
                out.extend(self.indent(body))
                first_clause = False
        self.expect("}")
        self.loop_stack.pop()
        if default_body is not None:
            if first_clause:
                out.extend(default_body)
            else:
                out.append("else:")
                out.extend(self.indent(default_body))
        self.scopes.pop()
        return out

    def _is_type_switch(self, q):
        depth = 0
        while True:
            tk = self.toks[q]
            if tk.k == "eof" or (tk.k == "op" and tk.v in ("{", ";") and depth == 0):
                return False
            if tk.k == "op" and tk.v in ("(", "["):
                if tk.v == "(" and self.toks[q - 1].k == "op" and self.toks[q - 1].v == "." and self.toks[q + 1].k == "kw" and self.toks[q + 1].v == "type":
                    return True
                depth += 1
            elif tk.k == "op" and tk.v in (")", "]"):
                depth -= 1
            q += 1

    def parse_primary_no_assert(self):
        """primary expression up to (not including) a trailing `.(type)`"""
        self.stop_at_type_assert = True
        try:
            return self.parse_unary()
        finally:
            self.stop_at_type_assert = False

    def type_object_code(self, ty):
        """python expression for a type used in assertions / type switches"""
        while ty[0] == "ptr":
            ty = ty[1]
        if ty[0] == "name":
            if ty[1] in self.types:
                return safe(ty[1])
            return repr(ty[1])
        if ty[0] == "qual":
            return "%s.%s" % (safe(ty[1]), safe(ty[2]))
        return repr(ty[0])

    # ------------------------------------------------------------------ expressions
    def parse_expr(self, prec=1):
        left = self.parse_unary()
        while True:
            t = self.peek()
            if t.k != "op" or t.v not in BINPREC or BINPREC[t.v] < prec:
                return left
            op = t.v
            self.next()
            right = self.parse_expr(BINPREC[op] + 1)
            left = self.binop(op, left, right, t)

    def fold(self, op, a, b):
        """exact value of `a op b` for two constants, or None"""
        (x, xf), (y, yf) = a, b
        f = xf or yf
        try:
            if op == "+":
                return (x + y, f)
            if op == "-":
                return (x - y, f)
            if op == "*":
                return (x * y, f)
            if op == "/":
                if y == 0:
                    return None
                if not f:
                    q = abs(x) // abs(y)
                    return (q if (x >= 0) == (y >= 0) else -q, False)
                return (Fraction(x) / Fraction(y), True)
            if f:
                return None
            if op == "%":
                r = abs(x) % abs(y)
                return (r if x >= 0 else -r, False)
            if op == "<<":
                return (x << y, False)
            if op == ">>":
                return (x >> y, False)
            if op == "&":
                return (x & y, False)
            if op == "|":
                return (x | y, False)
            if op == "^":
                return (x ^ y, False)
            if op == "&^":
                return (x & ~y, False)
        except (ZeroDivisionError, ValueError, OverflowError):
            return None
        return None

    def binop(self, op, a, b, t):
        if a.k is not None and b.k is not None and a.ityp is None and b.ityp is None:
            k = self.fold(op, a.k, b.k)
            if k is not None:
                r = Ex("(%s)" % kcode(k), "f" if k[1] else None, not k[1])
                r.k = k
                return r
        both_const = a.intconst and b.intconst
        if op in ("==", "!=", "<", "<=", ">", ">="):
            return Ex("(%s %s %s)" % (a.code, op, b.code), "b")
        if op == "&&":
            return Ex("(%s and %s)" % (a.code, b.code), "b")
        if op == "||":
            return Ex("(%s or %s)" % (a.code, b.code), "b")
        kind = None
        if a.kind == "c" or b.kind == "c":
            kind = "c"
        elif a.kind == "f" or b.kind == "f":
            kind = "f"
        elif a.kind == "s" or b.kind == "s":
            kind = "s"
        elif (a.kind == "i" or a.intconst) and (b.kind == "i" or b.intconst):
            kind = "i"
        if op == "/":
            if kind in ("f", "c"):
                code = "G.div(%s, %s)" % (a.code, b.code)
            else:
                code = "G.div(%s, %s, %r)" % (a.code, b.code, self.site(t))
        elif op == "%":
            code = "G.mod(%s, %s)" % (a.code, b.code)
        elif op == "&^":
            code = "G.andnot(%s, %s)" % (a.code, b.code)
        else:
            code = "(%s %s %s)" % (a.code, op, b.code)
        if op in ("<<", ">>"):
            kind = "i" if (a.kind == "i" or a.intconst) else kind
            both_const = a.intconst and b.intconst
        r = Ex(code, kind, both_const)
        ityp = a.ityp if op in ("<<", ">>") else (a.ityp or b.ityp)
        if ityp and kind in ("i", None) and not both_const:
            r.ityp = ityp
            if op in ("+", "-", "*", "<<"):
                r.code = "%s(%s)" % (INT_TYPES[ityp], code)
        return r

    def parse_unary(self):
        t = self.peek()
        if t.k == "op" and t.v in ("-", "+", "!", "^", "*", "&", "<-"):
            self.next()
            e = self.parse_unary()
            if t.v == "-" and e.k is not None:
                r = Ex("", "f" if e.k[1] else None, not e.k[1])
                r.k = (-e.k[0], e.k[1])
                r.code = "(%s)" % kcode(r.k)
                return r
            if t.v == "-":
                r = Ex("(-%s)" % e.code, e.kind, e.intconst)
                if e.ityp:
                    r.ityp = e.ityp
                    r.code = "%s(%s)" % (INT_TYPES[e.ityp], r.code)
                return r
            if t.v == "+":
                return e
            if t.v == "!":
                return Ex("(not %s)" % e.code, "b")
            if t.v == "^":
                return Ex("(~%s)" % e.code, "i", e.intconst)
            if t.v in ("*", "&"):
                return Ex(e.code, e.kind)                   # pointers to structs are the objects themselves
            return Ex("G.recv(%s)" % e.code)                 # <-ch
        return self.parse_primary()

    def parse_operand(self):
        t = self.next()
        if t.k == "int":
            txt = t.v.replace("_", "")
            if re.match(r"^0[0-7]+$", txt):
                val = int(txt, 8)
            else:
                val = int(txt, 0)
            e = Ex(str(val), None, True)
            e.k = (val, False)
            return e
        if t.k == "float":
            txt = t.v.replace("_", "")
            e = Ex(repr(float(txt)), "f")
            if not txt.lower().startswith("0x"):
                e.k = (Fraction(txt), True)
            return e
        if t.k == "imag":
            return Ex("complex(0.0, %r)" % float(t.v.replace("_", "")), "c")
        if t.k == "char":
            e = Ex(str(ord(t.v)), None, True)
            e.k = (ord(t.v), False)
            return e
        if t.k == "string":
            return Ex(repr(t.v), "s")
        if t.k == "op" and t.v == "(":
            nolit_save = self.nolit
            self.nolit = 0
            e = self.parse_expr()
            self.nolit = nolit_save
            self.expect(")")
            r = Ex("(%s)" % e.code if not e.code.startswith("(") else e.code, e.kind, e.intconst)
            r.k, r.ityp = e.k, e.ityp
            return r
        if t.k == "op" and t.v == "[":
            self.p -= 1
            ty = self.parse_type()
            if self.at("{"):
                return self.parse_composite(ty)
            if self.at("("):                                  # conversion []T(x)
                self.next(); e = self.parse_expr(); self.expect(")")
                return Ex(e.code)
            self.err("unexpected type in expression")
        if t.k == "kw" and t.v == "map":
            self.p -= 1
            ty = self.parse_type()
            if self.at("{"):
                return self.parse_composite(ty)
            self.err("unexpected map type in expression")
        if t.k == "kw" and t.v == "struct":
            self.err("anonymous struct literals are not supported")
        if t.k == "kw" and t.v == "func":
            self.err("function literals are not supported")
        if t.k == "ident":
            name = t.v
            loc = self.lookup(name) if hasattr(self, "scopes") and self.scopes else None
            if loc is not None:
                e = Ex(loc[0], loc[1])
                e.goname = name
                e.ityp = loc[2]
                return e
            if name == "true":
                return Ex("True", "b")
            if name == "false":
                return Ex("False", "b")
            if name == "nil":
                return Ex("None")
            if name == "iota":
                if getattr(self, "iota", None) is None:
                    self.err("iota outside a constant declaration")
                e = Ex(str(self.iota), None, True)
                e.k = (self.iota, False)
                return e
            e = Ex(safe(name))
            e.goname = name
            e.is_global = True
            if name in self.kconsts:
                e.k = self.kconsts[name]
                e.kind = "f" if e.k[1] else None
                e.intconst = not e.k[1]
                e.code = kcode(e.k)
            return e
        self.err("unexpected token %r in expression" % (t.v,), t)

    def parse_primary(self):
        t = self.peek()
        # builtin calls and conversions that need the type syntactically
        if t.k == "ident" and self.peek(1).k == "op" and self.peek(1).v == "(" and self.lookup(t.v) is None:
            name = t.v
            if name == "make":
                self.next(); self.next()
                ty = self.parse_type()
                args = []
                while self.accept(","):
                    args.append(self.parse_expr())
                self.expect(")")
                if ty[0] == "slice":
                    e = Ex("G.make_slice(%s, %s, %s, %s)" % (self.zero_code(ty[1], True), args[0].code, args[1].code if len(args) > 1 else "None", self.kind_code(ty[1])))
                elif ty[0] == "map":
                    e = Ex("{}")
                elif ty[0] == "chan":
                    e = Ex("G.Chan(%s)" % (args[0].code if args else "0"))
                else:
                    self.err("make of %r is not supported" % (ty[0],))
                return self.parse_postfix(e)
            if name == "new":
                self.next(); self.next()
                ty = self.parse_type(); self.expect(")")
                return self.parse_postfix(Ex(self.zero_code(ty)))
            if name in BASIC_CONV and name not in self.toplevel:
                self.next(); self.next()
                a = self.parse_expr(); self.expect(")")
                k = kind_of_type(("name", name))
                if name == "float64" and a.k is not None:
                    cv = Ex("", "f")
                    cv.k = (Fraction(float(a.k[0])), True)          # a typed constant: rounded to float64 here
                    cv.code = kcode(cv.k)
                    return self.parse_postfix(cv)
                if k == "f" and a.kind == "f":
                    return self.parse_postfix(Ex(a.code, "f"))
                cv = Ex("%s(%s)" % (BASIC_CONV[name], a.code), k)
                cv.ityp = sized_int(("name", name))
                return self.parse_postfix(cv)
            if name in self.types and isinstance(self.types[name], tuple):
                self.next(); self.next()
                a = self.parse_expr(); self.expect(")")
                return self.parse_postfix(Ex("%s(%s)" % (safe(name), a.code)))
        e = self.parse_operand()
        return self.parse_postfix(e)

    def parse_args(self):
        args, spread = [], False
        while not self.at(")"):
            args.append(self.parse_expr())
            if self.accept("..."):
                spread = True
            if not self.accept(","):
                break
        self.expect(")")
        return args, spread

    def parse_postfix(self, e):
        while True:
            t = self.peek()
            if t.k != "op":
                return e
            if t.v == ".":
                if self.peek(1).k == "op" and self.peek(1).v == "(":
                    if self.peek(2).k == "kw" and self.peek(2).v == "type":
                        return e                             # `.(type)`: handled by the switch
                    self.next(); self.next()
                    ty = self.parse_type(); self.expect(")")
                    tc = self.type_object_code(ty)
                    n = Ex("G.assert_type(%s, %s)" % (e.code, tc))
                    n.is_assert, n.assert_type, n.assert_operand = True, tc, e.code
                    e = n
                    continue
                self.next()
                name = self.ident()
                base_is_pkg = getattr(e, "goname", None) in self.imports and self.lookup(getattr(e, "goname", "")) is None
                if base_is_pkg and self.at("{") and not self.nolit:
                    e = self.parse_composite(("qual", e.goname, name))
                    continue
                n = Ex("%s.%s" % (e.code, safe(name)))
                if base_is_pkg:
                    n.qual = (e.goname, name)
                    kt = KCONST.get(self.imports[e.goname], {})
                    if name in kt:
                        n.k = kt[name]
                        n.kind = "f" if n.k[1] else None
                        n.intconst = not n.k[1]
                        n.code = kcode(n.k)
                e = n
                continue
            if t.v == "(":
                self.next()
                save_nolit = self.nolit
                self.nolit = 0
                callee = getattr(e, "goname", None)
                if callee in ("len", "cap", "append", "copy", "complex", "real", "imag", "panic", "delete", "min", "max", "print", "println", "close") \
                        and self.lookup(callee) is None and callee not in self.toplevel:
                    args, spread = self.parse_args()
                    self.nolit = save_nolit
                    e = self.builtin_call(callee, args, spread)
                    continue
                args, spread = self.parse_args()
                self.nolit = save_nolit
                parts = [a.code for a in args]
                if spread:
                    parts[-1] = "*G.spread(%s)" % parts[-1]
                n = Ex("%s(%s)" % (e.code, ", ".join(parts)))
                n.is_call = True
                e = n
                continue
            if t.v == "[":
                self.next()
                save_nolit = self.nolit
                self.nolit = 0
                lo = hi = mx = None
                is_slice = False
                if not self.at(":"):
                    lo = self.parse_expr()
                if self.accept(":"):
                    is_slice = True
                    if not self.at("]") and not self.at(":"):
                        hi = self.parse_expr()
                    if self.accept(":"):
                        mx = self.parse_expr()
                self.expect("]")
                self.nolit = save_nolit
                if is_slice:
                    e = Ex("G.slc(%s, %s, %s%s)" % (e.code, lo.code if lo else "None", hi.code if hi else "None", (", " + mx.code) if mx else ""))
                else:
                    n = Ex("%s[%s]" % (e.code, lo.code))
                    n.is_index, n.index_base, n.index_key, n.index_zero = True, e.code, lo.code, "None"
                    e = n
                continue
            if t.v == "{" and not self.nolit:
                nm = getattr(e, "goname", None)
                if nm and self.lookup(nm) is None and nm in self.types:
                    e = self.parse_composite(("name", nm))
                    continue
                return e
            return e

    def builtin_call(self, name, args, spread):
        c = [a.code for a in args]
        if name == "len":
            return Ex("G.len(%s)" % c[0], "i")
        if name == "cap":
            return Ex("G.cap(%s)" % c[0], "i")
        if name == "append":
            if spread:
                c[-1] = "*G.spread(%s)" % c[-1]
            return Ex("G.append(%s)" % ", ".join(c))
        if name == "copy":
            return Ex("G.copy(%s, %s)" % (c[0], c[1]), "i")
        if name == "complex":
            return Ex("G.cmplx(%s, %s)" % (c[0], c[1]), "c")
        if name == "real":
            return Ex("(%s).real" % c[0], "f")
        if name == "imag":
            return Ex("(%s).imag" % c[0], "f")
        if name == "panic":
            return Ex("G.panic(%s)" % c[0])
        if name == "delete":
            return Ex("%s.pop(%s, None)" % (c[0], c[1]))
        if name == "close":
            return Ex("G.close(%s)" % c[0])
        if name in ("min", "max"):
            return Ex("%s(%s)" % (name, ", ".join(c)), args[0].kind)
        return Ex("print(%s)" % ", ".join(c))

    def parse_composite(self, ty):
        """composite literal body `{...}` of type ty"""
        self.expect("{")
        save_nolit = self.nolit
        self.nolit = 0
        self.skip_semis()
        items = []                                          # (key Ex | None, value Ex)
        base = ty
        while base[0] == "ptr":
            base = base[1]
        elem_ty = base[2] if base[0] in ("array", "map") else (base[1] if base[0] == "slice" else None)
        is_struct = base[0] in ("name", "qual") and not (base[0] == "name" and isinstance(self.types.get(base[1]), tuple))

        def value():
            if self.at("{"):
                if elem_ty is None:
                    self.err("elided literal type without an element type")
                et = elem_ty
                while et[0] == "ptr":
                    et = et[1]
                return self.parse_composite(et)
            return self.parse_expr()
        while not self.at("}"):
            if is_struct and self.peek().k == "ident" and self.peek(1).k == "op" and self.peek(1).v == ":":
                k = Ex(self.ident()); self.next()
                items.append((k, value()))
            else:
                v = value()
                if self.accept(":"):
                    items.append((v, value()))
                else:
                    items.append((None, v))
            self.skip_semis()
            if not self.accept(","):
                self.skip_semis()
                break
            self.skip_semis()
        self.expect("}")
        self.nolit = save_nolit
        if is_struct:
            cls = safe(base[1]) if base[0] == "name" else "%s.%s" % (safe(base[1]), safe(base[2]))
            if items and items[0][0] is None:
                return Ex("%s._pos(%s)" % (cls, ", ".join(v.code for _, v in items)))
            return Ex("%s(%s)" % (cls, ", ".join("%s=%s" % (safe(k.code), v.code) for k, v in items)))
        if base[0] == "map":
            return Ex("{%s}" % ", ".join("%s: %s" % (k.code, v.code) for k, v in items))
        kind = self.kind_code(elem_ty)
        if any(k is not None for k, _ in items):
            pairs, nxt = [], 0
            for k, v in items:
                if k is not None:
                    pairs.append("(%s, %s)" % (k.code, v.code))
                else:
                    pairs.append("(%s, %s)" % ("_i", v.code))
                    self.err("mixed keyed / positional array literals are not supported")
            return Ex("G.lit_indexed([%s], %s, %s)" % (", ".join(pairs), self.zero_code(elem_ty, True), kind))
        code = "G.lit([%s], %s)" % (", ".join(v.code for _, v in items), kind)
        if base[0] == "array" and base[1] is not None:
            code = "G.lit_fixed(%s, %s, %s, %s)" % (code, base[1], self.zero_code(elem_ty, True), kind)
        return Ex(code)
