"""Loads Go packages of the reference tree as Python modules through go2py (test infrastructure; see go2py.py)."""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)
import go2py       # noqa: E402
import goruntime   # noqa: E402

MODULE_PREFIX = "github.com/andrepxx/go-dsp-guitar/"


class Missing(object):
    """an imported package that is not translated: any use raises"""

    def __init__(self, path):
        self._path = path

    def __getattr__(self, name):
        raise goruntime.GoPanic("package %s is not available in the translation (used: %s)" % (self._path, name))


class Loader(object):
    def __init__(self, root, overrides=None, skip=(), tolerant=()):
        self.root, self.mods, self.sources = root, {}, {}
        self.tolerant, self.skipped = set(tolerant), {}       # packages of which only part is translatable (e.g. controller)
        self.overrides = overrides or {}                    # import path -> python object standing in for the package
        self.skip = set(skip)

    def load(self, pkg):
        """pkg: directory name under the reference root (e.g. 'effects')"""
        if pkg in self.mods:
            return self.mods[pkg]
        d = os.path.join(self.root, pkg)
        files = []
        for fn in sorted(os.listdir(d)):
            if fn.endswith(".go") and not fn.endswith("_test.go"):
                files.append((os.path.join(d, fn), open(os.path.join(d, fn), encoding="utf-8").read()))
        tr = go2py.PackageTranslator(pkg, files, tolerant=pkg in self.tolerant).translate()
        self.skipped[pkg] = tr.skipped
        src, inits = tr.module_source()
        mod = types.ModuleType("goref_" + pkg.replace("/", "_"))
        g = mod.__dict__
        g["G"] = goruntime.G
        self.mods[pkg] = mod
        for name, path in tr.imports.items():
            if path in self.overrides:
                g[go2py.safe(name)] = self.overrides[path]
            elif path.startswith(MODULE_PREFIX):
                sub = path[len(MODULE_PREFIX):]
                g[go2py.safe(name)] = Missing(path) if sub in self.skip else self.load(sub)
            elif path in goruntime.STDLIB:
                g[go2py.safe(name)] = goruntime.STDLIB[path]
            else:
                g[go2py.safe(name)] = Missing(path)
        self.sources[pkg] = src + "\n\n# ---- package-level constants and variables\n" + "\n".join(inits)
        exec(compile(src, "<go2py:%s>" % pkg, "exec"), g)
        pending = list(inits)
        while pending:
            rest, progressed = [], False
            for line in pending:
                try:
                    exec(compile(line, "<go2py:%s:init>" % pkg, "exec"), g)
                    progressed = True
                except NameError:
                    rest.append(line)
            if not progressed:
                raise RuntimeError("package %s: unresolved package-level initialisers: %r" % (pkg, rest[:3]))
            pending = rest
        return mod
