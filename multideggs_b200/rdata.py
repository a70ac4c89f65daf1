"""Minimal reader for R `save()` files (RDX3 / XDR serialisation, gzip|xz|bzip2|raw).

There is no R in the build image or on the GPU box, so the bundled example data of the
reference (`data/*.rda`, `R/sysdata.rda`; documented in `R/example_data.R:1-40` and
`data-raw/omic_network.R:1-29`) is read with this file when the golden fixtures under
`tests/golden/` are generated (`tests/golden/make_golden.py`).  Only the SEXP types that
occur in numeric matrices, character vectors, factors and data.frames are handled.

Returned Python shapes:
  REALSXP/INTSXP/LGLSXP -> numpy array (column-major reshaped if a `dim` attribute exists)
  STRSXP                -> list[str|None]
  VECSXP                -> list
  attributes            -> `RObj.attr` dict (names, dim, dimnames, levels, class, row.names)
"""
from __future__ import annotations

import bz2
import gzip
import lzma
import struct
from typing import Any

import numpy as np

NA_INT = -2147483648


class RObj:
    """A decoded R value plus its attribute dict."""

    __slots__ = ("value", "attr")

    def __init__(self, value: Any, attr: dict | None = None):
        self.value = value
        self.attr = attr or {}

    def __repr__(self) -> str:  # pragma: no cover - debugging aid
        return f"RObj({type(self.value).__name__}, attr={list(self.attr)})"


def _decompress(raw: bytes) -> bytes:
    if raw[:2] == b"\x1f\x8b":
        return gzip.decompress(raw)
    if raw[:6] == b"\xfd7zXZ\x00":
        return lzma.decompress(raw)
    if raw[:3] == b"BZh":
        return bz2.decompress(raw)
    return raw


class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        self.p = 0
        self.refs: list[Any] = []

    def i32(self) -> int:
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def take(self, n: int) -> bytes:
        v = self.b[self.p:self.p + n]
        self.p += n
        return v

    def length(self) -> int:
        n = self.i32()
        if n == -1:  # long vector: two more ints
            hi, lo = self.i32(), self.i32()
            n = (hi << 32) + (lo & 0xFFFFFFFF)
        return n

    def item(self) -> Any:
        flags = self.i32()
        t = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        if t == 254:  # NILVALUE_SXP
            return None
        if t in (253, 252, 251, 250, 249, 248, 247, 242, 241):  # env/namespace pseudo-SEXPs
            if t in (249, 247):  # NAMESPACESXP / PACKAGESXP carry a STRSXP payload
                self.i32()
                n = self.i32()
                v = [self.item() for _ in range(n)]
                self.refs.append(v)
                return v
            return None
        if t == 255:  # REFSXP
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if t == 1:  # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if t in (2, 6):  # LISTSXP / LANGSXP  -> python list of (tag, value)
            out = []
            while True:
                attr = self.item() if has_attr else None  # noqa: F841 (pairlist attrs unused)
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                flags = self.i32()
                t2 = flags & 0xFF
                if t2 == 254:
                    break
                if t2 not in (2, 6):
                    raise ValueError(f"unexpected cdr type {t2}")
                has_attr = bool(flags & (1 << 9))
                has_tag = bool(flags & (1 << 10))
            return out
        if t == 9:  # CHARSXP
            n = self.i32()
            if n == -1:
                return None
            return self.take(n).decode("utf-8", errors="replace")
        if t == 238:  # ALTREP_SXP: (info, state, attr)
            info = self.item()
            state = self.item()
            attr = self.item()
            cls = info[0][1] if info else ""
            val = self._altrep(cls, state)
            return RObj(val, self._attrdict(attr)) if attr else val
        if t == 10 or t == 13:
            n = self.length()
            v = np.frombuffer(self.take(4 * n), dtype=">i4").astype(np.int32)
        elif t == 14:
            n = self.length()
            v = np.frombuffer(self.take(8 * n), dtype=">f8").astype(np.float64)
        elif t == 16:
            n = self.length()
            v = [self.item() for _ in range(n)]
        elif t in (19, 20):
            n = self.length()
            v = [self.item() for _ in range(n)]
        elif t == 24:  # RAWSXP
            n = self.length()
            v = self.take(n)
        else:
            raise NotImplementedError(f"SEXP type {t} at offset {self.p}")
        if has_attr:
            return RObj(v, self._attrdict(self.item()))
        return v

    @staticmethod
    def _attrdict(pl) -> dict:
        return {k: v for k, v in (pl or [])}

    @staticmethod
    def _altrep(cls: str, state: Any) -> Any:
        if cls in ("compact_intseq", "compact_realseq"):
            n, start, step = (float(x) for x in _val(state))
            seq = start + step * np.arange(int(n))
            return seq.astype(np.int32 if cls == "compact_intseq" else np.float64)
        if cls.startswith("wrap_"):
            return _val(state)[0] if isinstance(_val(state), list) else state
        if cls == "deferred_string":
            src = state[0][1] if isinstance(state, list) and isinstance(state[0], tuple) else state
            return [str(x) for x in np.asarray(_val(src)).tolist()]
        raise NotImplementedError(f"ALTREP class {cls}")


def _val(x: Any) -> Any:
    return x.value if isinstance(x, RObj) else x


def _attr(x: Any) -> dict:
    return x.attr if isinstance(x, RObj) else {}


def read_rda(path: str) -> dict[str, Any]:
    """Read an R `save()` file and return `{object name: decoded value}`."""
    with open(path, "rb") as fh:
        buf = _decompress(fh.read())
    if buf[:5] not in (b"RDX3\n", b"RDX2\n"):
        raise ValueError(f"{path}: not an RDX2/RDX3 file")
    r = _Reader(buf)
    r.p = 5
    if r.take(2) != b"X\n":
        raise ValueError("only XDR serialisation is supported")
    version = r.i32()
    r.i32()
    r.i32()
    if version == 3:
        r.take(r.i32())  # native encoding name
    top = r.item()
    return {tag: val for tag, val in top}


def as_matrix(x: Any) -> tuple[np.ndarray, list[str], list[str]]:
    """Decode an R numeric matrix -> (row-major ndarray[G, n], rownames, colnames)."""
    dim = _val(_attr(x)["dim"])
    g, n = int(dim[0]), int(dim[1])
    m = np.asarray(_val(x), dtype=np.float64).reshape((n, g)).T.copy()
    dn = _val(_attr(x).get("dimnames")) or [None, None]
    rn = list(_val(dn[0])) if dn[0] is not None else [str(i + 1) for i in range(g)]
    cn = list(_val(dn[1])) if dn[1] is not None else [str(i + 1) for i in range(n)]
    return m, rn, cn


def as_dataframe(x: Any) -> dict[str, Any]:
    """Decode an R data.frame -> dict(columns..., `_row_names`). Factors become list[str]."""
    a = _attr(x)
    names = list(_val(a["names"]))
    out: dict[str, Any] = {}
    for nm, col in zip(names, _val(x)):
        ca = _attr(col)
        if "levels" in ca:
            lev = list(_val(ca["levels"]))
            codes = np.asarray(_val(col))
            out[nm] = [None if c == NA_INT else lev[c - 1] for c in codes]
            out[nm + "._levels"] = lev
        else:
            out[nm] = _val(col)
    rn = _val(a.get("row.names"))
    if isinstance(rn, np.ndarray) and rn.size == 2 and rn[0] == NA_INT:
        rn = [str(i + 1) for i in range(abs(int(rn[1])))]
    elif isinstance(rn, np.ndarray):
        rn = [str(i) for i in rn.tolist()]
    out["_row_names"] = list(rn) if rn is not None else None
    return out
