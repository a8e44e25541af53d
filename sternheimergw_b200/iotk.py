"""iotk *binary* XML container (the QE I/O toolkit the reference links): reader and writer for the two files of the path that use it --

* the self-energy file of ``sigma_io_module`` (algo/io/src/sigma_io.f90:70-204: root SELF_ENERGY, NUM_EXCHANGE, NUM_CORRELATION,
  NUM_FREQUENCY, KPOINT, then per k-point <SIGMA.ik> CORRELATION, EXCHANGE </SIGMA.ik>), and
* the solver fixture ``green_solver_debug`` writes (phys/green/src/green.f90:438-446: root LINEAR_PROBLEM, DIMENSION, NUMBER_SHIFT,
  LIST_SHIFT, LINEAR_OPERATOR, RIGHT_HAND_SIDE, INCORRECT_SOLUTION) -- the reference's own golden vector
  algo/linear_solver/test/lin_prob.xml.bz2 is such a file.

The iotk source is not vendored in the reference tree; the layout below was read off that golden file record by record and is
pinned by regenerating it BYTE FOR BYTE (tests/test_iotk.py).  A binary iotk file is a sequence of Fortran unformatted records
(little-endian int32 length, payload, int32 length):

    tag            two records:  [int32 256 * len(text) + control]   and   [int32 256 * len(text) + 128][text]
                   control 1 = begin, 2 = end, 3 = empty, 5 = processing instruction; text = "\\n" + indent + "<...>" + "\\n",
                   two blanks of indent per nesting level
    data           one record:   [int32 0][raw array, Fortran element order]      between <NAME type= size= kind=> and </NAME>
                   type integer / real / complex, kind = bytes of one (real) component, size = number of elements

Host-side I/O (numpy); nothing here touches the device.
"""
from __future__ import annotations

import struct
from pathlib import Path

import numpy as np

CONTROL_BEGIN, CONTROL_END, CONTROL_EMPTY, CONTROL_PI = 1, 2, 3, 5
_HEADER_PI = ('iotk version="1.2.0"', 'iotk file_version="1.0"', 'iotk binary="T"', 'iotk qe_syntax="F"')


def _record(payload: bytes) -> bytes:
    n = struct.pack("<i", len(payload))
    return n + payload + n


def iotk_index(i: int) -> str:
    """iotk_index(i): the suffix that makes a tag unique (sigma_io.f90:170)."""
    return f".{int(i)}"


def _typeinfo(a: np.ndarray):
    if np.issubdtype(a.dtype, np.complexfloating):
        return "complex", 8, "<c16"
    if np.issubdtype(a.dtype, np.floating):
        return "real", 8, "<f8"
    if np.issubdtype(a.dtype, np.integer):
        return "integer", 4, "<i4"
    raise TypeError(f"iotk: unsupported dtype {a.dtype}")


class IotkBinaryWriter:
    """iotk_open_write(unit, file, binary=.TRUE., root=...) ... iotk_close_write(unit)."""

    def __init__(self, path, root: str):
        self._f = open(path, "wb")
        self._level = 0
        self._root = root
        for pi in _HEADER_PI:
            self._tag(CONTROL_PI, f"<?{pi}?>")
        self.write_begin(root)

    def _tag(self, control: int, body: str):
        text = ("\n" + "  " * self._level + body + "\n").encode("ascii")
        self._f.write(_record(struct.pack("<i", 256 * len(text) + control)))
        self._f.write(_record(struct.pack("<i", 256 * len(text) + 128) + text))

    def write_begin(self, name: str, attrs: str = ""):
        self._tag(CONTROL_BEGIN, f"<{name}{(' ' + attrs) if attrs else ''}>")
        self._level += 1

    def write_end(self, name: str):
        self._level -= 1
        self._tag(CONTROL_END, f"</{name}>")

    def write_dat(self, name: str, dat):
        """iotk_write_dat: scalars and arrays of integer / real(dp) / complex(dp); arrays go out in Fortran element order."""
        a = np.asarray(dat)
        typ, kind, dt = _typeinfo(a)
        flat = np.asarray(a, dtype=dt).reshape(-1, order="F")
        self.write_begin(name, f'type="{typ}" size="{flat.size}" kind="{kind}"')
        self._f.write(_record(struct.pack("<i", 0) + flat.tobytes()))
        self.write_end(name)

    def close(self):
        self.write_end(self._root)
        self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def iotk_records(buf: bytes):
    """Yield (control, text | None, data bytes | None) for every logical item of a binary iotk file."""
    off, n = 0, len(buf)
    while off < n:
        (ln,) = struct.unpack_from("<i", buf, off)
        pay = buf[off + 4:off + 4 + ln]
        (ln2,) = struct.unpack_from("<i", buf, off + 4 + ln)
        if ln != ln2:
            raise ValueError("iotk: corrupt Fortran record")
        off += 8 + ln
        (hdr,) = struct.unpack_from("<i", pay, 0)
        if ln == 4:                                    # first half of a tag: control word; the text follows in the next record
            control = hdr % 256
            (l2,) = struct.unpack_from("<i", buf, off)
            pay2 = buf[off + 4:off + 4 + l2]
            off += 8 + l2
            (h2,) = struct.unpack_from("<i", pay2, 0)
            if h2 != (hdr - control) + 128 or len(pay2) - 4 != hdr // 256:
                raise ValueError("iotk: tag header mismatch")
            yield control, pay2[4:].decode("ascii"), None
        else:
            if hdr != 0:
                raise ValueError("iotk: unexpected data header")
            yield 0, None, pay[4:]


class IotkBinaryReader:
    """iotk_open_read(binary=.TRUE.) + iotk_scan_begin / iotk_scan_dat / iotk_scan_end on an in-memory copy of the file."""

    def __init__(self, source):
        buf = source if isinstance(source, (bytes, bytearray)) else Path(source).read_bytes()
        self.items = list(iotk_records(bytes(buf)))
        self.root = None
        self._tree = self._build()

    @staticmethod
    def _parse_tag(text: str):
        body = text.strip()
        inner = body.strip("<>/?").strip()
        parts = inner.split(None, 1)
        name = parts[0]
        attrs = {}
        if len(parts) > 1:
            import re
            attrs = dict(re.findall(r'(\w+)="([^"]*)"', parts[1]))
        return name, attrs

    def _build(self):
        root = {"name": None, "attrs": {}, "children": [], "data": None}
        stack = [root]
        for control, text, data in self.items:
            if control == CONTROL_PI:
                continue
            if control == CONTROL_BEGIN:
                name, attrs = self._parse_tag(text)
                node = {"name": name, "attrs": attrs, "children": [], "data": None}
                stack[-1]["children"].append(node)
                stack.append(node)
                if self.root is None:
                    self.root = name
            elif control == CONTROL_END:
                name, _ = self._parse_tag(text)
                if stack[-1]["name"] != name:
                    raise ValueError(f"iotk: </{name}> closes <{stack[-1]['name']}>")
                stack.pop()
            elif control == CONTROL_EMPTY:
                name, attrs = self._parse_tag(text)
                stack[-1]["children"].append({"name": name, "attrs": attrs, "children": [], "data": None})
            elif control == 0:
                stack[-1]["data"] = data
        if len(stack) != 1:
            raise ValueError("iotk: unbalanced tags")
        return root["children"][0]

    def _find(self, path):
        node = self._tree
        for name in path:
            hits = [c for c in node["children"] if c["name"] == name]
            if not hits:
                raise KeyError("/".join(path))
            node = hits[0]
        return node

    def names(self, *path):
        return [c["name"] for c in self._find(path)["children"]]

    def scan_dat(self, *path, shape=None):
        """iotk_scan_dat: the array below root/path...; `shape` (Fortran order) reshapes it as the caller's array would."""
        node = self._find(path)
        typ, size, kind = node["attrs"]["type"], int(node["attrs"]["size"]), int(node["attrs"]["kind"])
        dt = {"integer": f"<i{kind}", "real": f"<f{kind}", "complex": f"<c{2 * kind}"}[typ]
        a = np.frombuffer(node["data"], dtype=dt, count=size).copy()
        if shape is not None:
            a = a.reshape(shape, order="F")
        return a


# ----------------------------------------------------------------------------- sigma_io_module (algo/io/src/sigma_io.f90)
TAG_ROOT, TAG_NUM_EXCHANGE, TAG_NUM_CORRELATION, TAG_FREQUENCY = "SELF_ENERGY", "NUM_EXCHANGE", "NUM_CORRELATION", "NUM_FREQUENCY"
TAG_KPOINT, TAG_SIGMA, TAG_EXCHANGE, TAG_CORRELATION = "KPOINT", "SIGMA", "EXCHANGE", "CORRELATION"


def sigma_io_open_write(filename, kpt, ngm_x: int, ngm_c: int, num_freq: int) -> IotkBinaryWriter:
    """sigma_io_open_write (sigma_io.f90:70): header with the metadata; kpt is (3, nks) as in the reference."""
    w = IotkBinaryWriter(filename, TAG_ROOT)
    w.write_dat(TAG_NUM_EXCHANGE, np.int32(ngm_x))
    w.write_dat(TAG_NUM_CORRELATION, np.int32(ngm_c))
    w.write_dat(TAG_FREQUENCY, np.int32(num_freq))
    w.write_dat(TAG_KPOINT, np.asarray(kpt, dtype=np.float64))
    return w


def sigma_io_write_c(w: IotkBinaryWriter, ikpt: int, sigma_c):
    """sigma_io_write_c (sigma_io.f90:151): opens <SIGMA.ikpt> and writes the correlation part (must come first)."""
    w.write_begin(TAG_SIGMA + iotk_index(ikpt))
    w.write_dat(TAG_CORRELATION, np.asarray(sigma_c, dtype=np.complex128))


def sigma_io_write_x(w: IotkBinaryWriter, ikpt: int, sigma_x):
    """sigma_io_write_x (sigma_io.f90:185): writes the exchange part and closes <SIGMA.ikpt>."""
    w.write_dat(TAG_EXCHANGE, np.asarray(sigma_x, dtype=np.complex128))
    w.write_end(TAG_SIGMA + iotk_index(ikpt))


def sigma_io_close_write(w: IotkBinaryWriter):
    w.close()


def sigma_io_open_read(filename):
    """sigma_io_open_read (sigma_io.f90:113): returns (reader, kpt(3, nks), ngm_x, ngm_c, num_freq)."""
    r = IotkBinaryReader(filename)
    if r.root != TAG_ROOT:
        raise ValueError(f"not a self-energy file (root {r.root})")
    ngm_x = int(r.scan_dat(TAG_NUM_EXCHANGE)[0])
    ngm_c = int(r.scan_dat(TAG_NUM_CORRELATION)[0])
    num_freq = int(r.scan_dat(TAG_FREQUENCY)[0])
    kpt = r.scan_dat(TAG_KPOINT)
    return r, kpt.reshape((3, kpt.size // 3), order="F"), ngm_x, ngm_c, num_freq


def sigma_io_read(r: IotkBinaryReader, ikpt: int, ngm_x: int, ngm_c: int, num_freq: int):
    """sigma_io_read (sigma_io.f90:216): (sigma_x(ngm_x, ngm_x), sigma_c(ngm_c, ngm_c, num_freq)) of k-point ikpt."""
    tag = TAG_SIGMA + iotk_index(ikpt)
    sigma_c = r.scan_dat(tag, TAG_CORRELATION, shape=(ngm_c, ngm_c, num_freq))
    sigma_x = r.scan_dat(tag, TAG_EXCHANGE, shape=(ngm_x, ngm_x))
    return sigma_x, sigma_c


# ----------------------------------------------------------------------------- green_solver_debug (phys/green/src/green.f90:438-446)
def linear_problem_write(filename, hamil, bb, omega, green):
    """The file green_solver_debug dumps for the solver unit test: root LINEAR_PROBLEM."""
    hamil = np.asarray(hamil, dtype=np.complex128)
    with IotkBinaryWriter(filename, "LINEAR_PROBLEM") as w:
        w.write_dat("DIMENSION", np.int32(hamil.shape[0]))
        w.write_dat("NUMBER_SHIFT", np.int32(np.asarray(omega).size))
        w.write_dat("LIST_SHIFT", np.asarray(omega, dtype=np.complex128))
        w.write_dat("LINEAR_OPERATOR", hamil)
        w.write_dat("RIGHT_HAND_SIDE", np.asarray(bb, dtype=np.complex128))
        w.write_dat("INCORRECT_SOLUTION", np.asarray(green, dtype=np.complex128))


def linear_problem_read(source):
    """linear_problem_read of algo/linear_solver/test/linear_solver.pf:44-97: (A, b, sigma, x_bad)."""
    r = IotkBinaryReader(source)
    if r.root != "LINEAR_PROBLEM":
        raise ValueError(f"not a LINEAR_PROBLEM file (root {r.root})")
    n = int(r.scan_dat("DIMENSION")[0])
    ns = int(r.scan_dat("NUMBER_SHIFT")[0])
    sigma = r.scan_dat("LIST_SHIFT")
    a = r.scan_dat("LINEAR_OPERATOR", shape=(n, n))
    b = r.scan_dat("RIGHT_HAND_SIDE")
    x = r.scan_dat("INCORRECT_SOLUTION", shape=(n, ns))
    return a, b, sigma, x
