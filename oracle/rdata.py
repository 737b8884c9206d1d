"""Minimal reader for R's RDX2/XDR serialisation, enough for bWGR's data/tpod.RData.

TEST INFRASTRUCTURE ONLY (see oracle/bwgr_oracle.hpp).  Used once, in this container, by
oracle/make_golden.py to turn /root/reference/data/tpod.RData (man/tpod.Rd:17) into
tests/golden/tpod.npz; nothing reads /root/reference at test or bench time.

Format: xz container -> "RDX2\\n" + "X\\n" + 3 int32 versions + one pairlist.
SEXP header int: type = low 8 bits, bit 8 object, bit 9 attributes, bit 10 tag.
"""
import lzma
import struct

import numpy as np


class _Reader:
    def __init__(self, buf):
        self.buf = buf
        self.pos = 0
        self.refs = []

    def i32(self):
        (v,) = struct.unpack_from(">i", self.buf, self.pos)
        self.pos += 4
        return v

    def item(self):
        flags = self.i32()
        typ = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        if typ == 254:  # NILVALUE
            return None
        if typ == 255:  # REFSXP
            return self.refs[(flags >> 8) - 1]
        if typ == 1:  # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if typ == 2:  # LISTSXP
            out = []
            while True:
                attr = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                del attr
                flags = self.i32()
                typ = flags & 0xFF
                if typ == 254:
                    break
                if typ != 2:
                    raise ValueError("unexpected cdr type %d" % typ)
                has_attr = bool(flags & (1 << 9))
                has_tag = bool(flags & (1 << 10))
            return out
        if typ == 9:  # CHARSXP
            n = self.i32()
            if n == -1:
                return None
            s = self.buf[self.pos:self.pos + n].decode("latin-1")
            self.pos += n
            return s
        if typ in (10, 13):  # LGLSXP, INTSXP
            n = self.i32()
            v = np.frombuffer(self.buf, dtype=">i4", count=n, offset=self.pos).astype(np.int32)
            self.pos += 4 * n
            return self._with_attr(v, has_attr)
        if typ == 14:  # REALSXP
            n = self.i32()
            v = np.frombuffer(self.buf, dtype=">f8", count=n, offset=self.pos).astype(np.float64)
            self.pos += 8 * n
            return self._with_attr(v, has_attr)
        if typ == 16:  # STRSXP
            n = self.i32()
            v = [self.item() for _ in range(n)]
            return self._with_attr(v, has_attr)
        if typ == 19:  # VECSXP
            n = self.i32()
            v = [self.item() for _ in range(n)]
            return self._with_attr(v, has_attr)
        raise ValueError("unsupported SEXP type %d at %d" % (typ, self.pos))

    def _with_attr(self, v, has_attr):
        if not has_attr:
            return v
        attrs = dict(self.item())
        if isinstance(v, np.ndarray) and "dim" in attrs:
            v = v.reshape(tuple(int(x) for x in attrs["dim"]), order="F")
        return v


def read_rdata(path):
    raw = open(path, "rb").read()
    if raw[:6] == b"\xfd7zXZ\x00":
        raw = lzma.decompress(raw)
    elif raw[:2] == b"\x1f\x8b":
        import gzip
        raw = gzip.decompress(raw)
    if raw[:5] != b"RDX2\n" or raw[5:7] != b"X\n":
        raise ValueError("not an RDX2/XDR file")
    rd = _Reader(raw)
    rd.pos = 7
    rd.i32(); rd.i32(); rd.i32()
    return dict(rd.item())


if __name__ == "__main__":
    import sys
    d = read_rdata(sys.argv[1])
    for k, v in d.items():
        print(k, getattr(v, "shape", None) or len(v))
