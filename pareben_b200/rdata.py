"""Reader for R's XDR serialisation (.rda / .RDS), no R needed: the wire format of the reference's bundled inputs
(data/BASIS.rda, y.rda, BASISbinomial.rda, yBinomial.rda, see /root/reference/EBEN_orig/man/BASIS.Rd) and of the saved
CrossValidate outputs under paper_materials/.  Handles the SEXP types that occur in those files (vectors, lists,
pairlists, attributes incl. dim/names/class, reference objects).  Used by pareben_b200.io / the command line, and by the
test infrastructure (oracle/rdata.py re-exports it).
"""
from __future__ import annotations

import bz2
import gzip
import lzma
import struct

import numpy as np

NA_INT = -2147483648


class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        self.p = 0
        self.refs: list = []

    def take(self, n: int) -> bytes:
        out = self.b[self.p:self.p + n]
        if len(out) != n:
            raise ValueError("truncated R data stream")
        self.p += n
        return out

    def i32(self) -> int:
        return struct.unpack(">i", self.take(4))[0]

    def length(self) -> int:
        n = self.i32()
        if n == -1:  # long vector
            hi, lo = self.i32(), self.i32()
            n = (hi << 32) + lo
        return n

    def item(self):
        flags = self.i32()
        t = flags & 0xFF
        has_attr = bool(flags & 0x200)
        has_tag = bool(flags & 0x400)
        if t == 254:   # NILVALUE_SXP
            return None
        if t == 253 or t == 242 or t == 241:  # global / empty / base env
            return {"env": t}
        if t == 255:   # REFSXP
            return self.refs[(flags >> 8) - 1]
        if t == 1:     # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if t == 2:     # LISTSXP (pairlist)
            out = []
            while True:
                attr = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                nxt = self.i32()
                nt = nxt & 0xFF
                if nt == 254:
                    break
                if nt != 2:
                    raise ValueError(f"unexpected pairlist tail type {nt}")
                has_attr = bool(nxt & 0x200)
                has_tag = bool(nxt & 0x400)
            return out
        if t == 9:     # CHARSXP
            n = self.i32()
            return None if n == -1 else self.take(n).decode("utf-8", "replace")
        if t == 10 or t == 13:   # LGLSXP / INTSXP
            n = self.length()
            val = np.frombuffer(self.take(4 * n), dtype=">i4").astype(np.int32)
        elif t == 14:  # REALSXP
            n = self.length()
            val = np.frombuffer(self.take(8 * n), dtype=">f8").astype(np.float64)
        elif t == 16:  # STRSXP
            n = self.length()
            val = [self.item() for _ in range(n)]
        elif t == 19 or t == 20:  # VECSXP / EXPRSXP
            n = self.length()
            val = [self.item() for _ in range(n)]
        else:
            raise ValueError(f"unsupported SEXP type {t} at byte {self.p}")
        attrs = {}
        if has_attr:
            for tag, car in (self.item() or []):
                attrs[tag] = car
        return _finish(val, attrs)


class RList(list):
    """An R list / data.frame: a Python list with .names and dict-style access by name."""

    names: list | None = None
    attrs: dict

    def __getitem__(self, k):
        if isinstance(k, str):
            return list.__getitem__(self, self.names.index(k))
        return list.__getitem__(self, k)

    def keys(self):
        return list(self.names or [])


def _finish(val, attrs):
    if isinstance(val, np.ndarray):
        dim = attrs.get("dim")
        if dim is not None:
            val = val.reshape(tuple(int(d) for d in dim), order="F")
        return val
    if isinstance(val, list) and (attrs.get("names") is not None or "class" in attrs):
        out = RList(val)
        out.names = list(attrs["names"]) if attrs.get("names") is not None else None
        out.attrs = attrs
        return out
    return val


def _decompress(raw: bytes) -> bytes:
    if raw[:2] == b"\x1f\x8b":
        return gzip.decompress(raw)
    if raw[:6] == b"\xfd7zXZ\x00":
        return lzma.decompress(raw)
    if raw[:3] == b"BZh":
        return bz2.decompress(raw)
    return raw


def _stream(buf: bytes) -> _Reader:
    if buf[:2] != b"X\n":
        raise ValueError("only XDR-format R serialisation is supported")
    r = _Reader(buf[2:])
    r.i32(); r.i32(); r.i32()          # format version, writer R version, min reader version
    return r


def read_rds(path: str):
    buf = _decompress(open(path, "rb").read())
    r = _stream(buf)
    return r.item()


def read_rda(path: str) -> dict:
    """Return {object name: value} for an .rda / .RData file."""
    buf = _decompress(open(path, "rb").read())
    if buf[:5] != b"RDX2\n":
        raise ValueError("not an RDX2 file")
    r = _stream(buf[5:])
    top = r.item()
    return {tag: val for tag, val in top}
