"""Python stand-ins for the few things the translated packages reach for outside the translated subset (test infrastructure):
os.ReadFile, encoding/json.Unmarshal (fills the slice it is handed, like json does through the pointer) and the RIFF
container reader of wave.FromBuffer (encoding/binary + reflection; SURVEY.md 8 puts the container out of scope).  The sample
conversion inside is wave.go's own rule for 16-bit PCM: float = int16 * 2 / (MaxInt16 - MinInt16) (wave/wave.go:399-427)."""
import json
import os
import struct

import goruntime
from goruntime import G


class OsShim(object):
    def __init__(self, root):
        self.root = root

    def ReadFile(self, path):
        p = path if os.path.isabs(path) else os.path.join(self.root, path)
        try:
            with open(p, "rb") as f:
                return G.lit(list(f.read()), None), None
        except OSError as e:
            return None, goruntime.GoError(str(e))


class JsonShim(object):
    elem_class = None            # the struct type of the list being decoded (set by the caller)

    def Unmarshal(self, content, target):
        try:
            data = json.loads(bytes(content.tolist()).decode("utf-8"))
        except ValueError as e:
            return goruntime.GoError(str(e))
        items = []
        for d in data:
            o = self.elem_class()
            for k, v in d.items():
                setattr(o, k, v)
            items.append(o)
        target.a, target.o, target.n, target.c = items, 0, len(items), len(items)
        return None


class _Channel(object):
    def __init__(self, floats):
        self._f = floats

    def Floats(self):
        return G.lit(list(self._f), 'f')


class _WaveFile(object):
    def __init__(self, rate, channels):
        self._rate, self._ch = rate, channels

    def ChannelCount(self):
        return len(self._ch)

    def SampleRate(self):
        return self._rate

    def Channel(self, i):
        if 0 <= i < len(self._ch):
            return _Channel(self._ch[i]), None
        return None, goruntime.GoError("no channel %d" % i)


class WaveShim(object):
    def FromBuffer(self, buf):
        b = bytes(buf.tolist())
        if b[:4] != b"RIFF" or b[8:12] != b"WAVE":
            return None, goruntime.GoError("not a RIFF/WAVE buffer")
        pos, fmt, data = 12, None, None
        while pos + 8 <= len(b):
            cid, size = b[pos:pos + 4], struct.unpack("<I", b[pos + 4:pos + 8])[0]
            body = b[pos + 8:pos + 8 + size]
            if cid == b"fmt ":
                fmt = struct.unpack("<HHIIHH", body[:16])
            elif cid == b"data":
                data = body
            pos += 8 + size + (size & 1)
        if fmt is None or data is None:
            return None, goruntime.GoError("missing chunk")
        tag, nch, rate, _, _, bits = fmt
        if tag != 1 or bits != 16:
            return None, goruntime.GoError("the shim reads 16-bit PCM only")
        ints = struct.unpack("<%dh" % (len(data) // 2), data[:2 * (len(data) // 2)])
        scaling = 2.0 / (32767 - (-32768))                 # wave.go:414
        chans = [[scaling * float(v) for v in ints[c::nch]] for c in range(nch)]
        return _WaveFile(rate, chans), None
