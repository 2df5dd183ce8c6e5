"""Minimal PLY reader / writer for 3DGS point clouds.

The reference uses the third-party `plyfile` package (not installed here, no network):
`PlyData.read` (deep_learning_segmentation.py:29, k_means.py:206), `PlyElement.describe` +
`PlyData([...], text=False).write` (deep_learning_segmentation.py:331-332, binary
little-endian) and `text=True` (k_means.py:190-193, ASCII).  This module produces the same
files: header `ply / format ... 1.0 / element vertex N / property <type> <name> ... /
end_header`, packed little-endian records, and for ASCII one vertex per line with every field
printed through `%.18g` of its float64 value (plyfile's `_write_txt`).

Only what the labelled-PLY path needs is supported: scalar properties (list properties are
rejected), any number of elements (non-vertex elements are carried through verbatim for
binary files).  plyfile itself is not available here; tests/test_ply_bytes.py pins the written
bytes from the consumer's side instead (a restatement of the viewer's parser,
Web_Viewer_Gaussians_Selection/gaussians_selection.js:464-511 and :579, and literal known answers).
"""
from __future__ import annotations

import numpy as np

# PLY type name -> numpy code, both spellings plyfile accepts on read.
_PLY_TO_NP = {
    "char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1",
    "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
    "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4",
    "float": "f4", "float32": "f4", "double": "f8", "float64": "f8",
}
# numpy code -> the name plyfile writes.
_NP_TO_PLY = {"i1": "char", "u1": "uchar", "i2": "short", "u2": "ushort",
              "i4": "int", "u4": "uint", "f4": "float", "f8": "double"}


class PlyElementData:
    """One element: `.name`, `.data` (structured array).  Indexing by property name works
    like plyfile's PlyElement (`vertices['x']`), and len() is the element count."""

    def __init__(self, name: str, data: np.ndarray):
        self.name = name
        self.data = data

    def __getitem__(self, key):
        return self.data[key]

    def __setitem__(self, key, value):
        self.data[key] = value

    def __len__(self):
        return len(self.data)


class PlyFile:
    """Parsed PLY: `.elements` in file order, `ply['vertex']`, `.text`, `.comments`."""

    def __init__(self, elements, text=False, comments=None):
        self.elements = list(elements)
        self.text = text
        self.comments = list(comments or [])

    def __getitem__(self, name):
        for e in self.elements:
            if e.name == name:
                return e
        raise KeyError(name)

    def write(self, path_or_file):
        write_ply(path_or_file, self.elements, text=self.text, comments=self.comments)


def _parse_header(f):
    if f.readline().strip() != b"ply":
        raise ValueError("not a PLY file")
    fmt, comments, elements = None, [], []
    while True:
        line = f.readline()
        if not line:
            raise ValueError("PLY header not terminated")
        tok = line.decode("ascii", "replace").strip().split()
        if not tok:
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "comment":
            comments.append(line.decode("ascii", "replace").strip()[8:])
        elif tok[0] == "element":
            elements.append((tok[1], int(tok[2]), []))
        elif tok[0] == "property":
            if tok[1] == "list":
                raise ValueError("list properties are not supported by this reader")
            elements[-1][2].append((tok[2], _PLY_TO_NP[tok[1]]))
        elif tok[0] == "end_header":
            break
    if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
        raise ValueError(f"unsupported PLY format {fmt!r}")
    return fmt, comments, elements


def read_ply(path_or_file) -> PlyFile:
    own = isinstance(path_or_file, (str, bytes)) or hasattr(path_or_file, "__fspath__")
    f = open(path_or_file, "rb") if own else path_or_file
    try:
        fmt, comments, header = _parse_header(f)
        out = []
        for name, count, props in header:
            if fmt == "ascii":
                dt = np.dtype([(n, "<" + c) for n, c in props])
                arr = np.empty(count, dt)
                if count:
                    # one vertex per line, every field a decimal number: NumPy's C text reader
                    # parses the whole element at once (float64, the precision plyfile prints with)
                    cols = np.loadtxt(f, dtype=np.float64, max_rows=count, ndmin=2)
                    if cols.shape != (count, len(props)):
                        raise ValueError(f"PLY element {name!r}: expected {count} x {len(props)} values, found {cols.shape}")
                    for j, (n, c) in enumerate(props):
                        arr[n] = cols[:, j] if c[0] == "f" else cols[:, j].astype(np.int64)
            else:
                order = "<" if fmt == "binary_little_endian" else ">"
                dt = np.dtype([(n, order + c) for n, c in props])
                raw = f.read(dt.itemsize * count)
                if len(raw) != dt.itemsize * count:
                    raise ValueError(f"PLY element {name!r} truncated")
                arr = np.frombuffer(raw, dt, count).astype(dt.newbyteorder("<")).copy()
            out.append(PlyElementData(name, arr))
        return PlyFile(out, text=(fmt == "ascii"), comments=comments)
    finally:
        if own:
            f.close()


def _header(elements, text, comments):
    lines = ["ply", "format ascii 1.0" if text else "format binary_little_endian 1.0"]
    lines += ["comment " + c for c in comments]
    for e in elements:
        lines.append(f"element {e.name} {len(e.data)}")
        for n in e.data.dtype.names:
            code = e.data.dtype[n].str.lstrip("<>|=")
            lines.append(f"property {_NP_TO_PLY[code]} {n}")
    lines.append("end_header")
    return ("\n".join(lines) + "\n").encode("ascii")


def write_ply(path_or_file, elements, text=False, comments=()):
    """Write elements (PlyElementData or (name, structured array) pairs)."""
    els = [e if isinstance(e, PlyElementData) else PlyElementData(*e) for e in elements]
    own = isinstance(path_or_file, (str, bytes)) or hasattr(path_or_file, "__fspath__")
    f = open(path_or_file, "wb") if own else path_or_file
    try:
        f.write(_header(els, text, comments))
        for e in els:
            names = e.data.dtype.names
            if text:
                # plyfile: every field of a record goes through '%.18g' as a float64
                if not _write_ascii_native(f, e.data):
                    cols = np.column_stack([e.data[n].astype(np.float64) for n in names]) if len(e.data) else np.empty((0, len(names)))
                    chunk = 65536
                    for s in range(0, len(cols), chunk):
                        np.savetxt(f, cols[s:s + chunk], fmt="%.18g", newline="\n")
            else:
                packed = np.dtype([(n, "<" + e.data.dtype[n].str.lstrip("<>|=")) for n in names])
                f.write(np.ascontiguousarray(e.data.astype(packed)).tobytes())
    finally:
        if own:
            f.close()


_TYPE_CODE = {"i1": 0, "u1": 1, "i2": 2, "u2": 3, "i4": 4, "u4": 5, "f4": 6, "f8": 7}


def _write_ascii_native(f, data) -> bool:
    """ASCII rows through the multi-threaded C formatter of libgslift.so (gsl_ply_format_ascii).
    Returns False when the library is not built, so that file I/O keeps working without it (this
    is host-side formatting, not part of the GPU path)."""
    import ctypes
    try:
        from ._native import lib
        L = lib()
    except (ImportError, OSError):
        return False
    names = data.dtype.names
    packed = np.dtype([(n, "<" + data.dtype[n].str.lstrip("<>|=")) for n in names])
    types = np.array([_TYPE_CODE[packed[n].str.lstrip("<>|=")] for n in names], np.int32)
    offsets = np.array([packed.fields[n][1] for n in names], np.int32)
    chunk = 1 << 15
    out = ctypes.create_string_buffer(chunk * len(names) * 40)
    for s in range(0, len(data), chunk):
        rec = np.ascontiguousarray(data[s:s + chunk].astype(packed))
        n = L.gsl_ply_format_ascii(rec.ctypes.data, len(rec), packed.itemsize, len(names), types.ctypes.data,
                                   offsets.ctypes.data, out, len(out), 0)
        if n < 0:
            raise RuntimeError(L.gsl_last_error().decode())
        f.write(memoryview(out)[:n])
    return True


def describe_with_label(vertex: np.ndarray, labels, name="label") -> np.ndarray:
    """Vertex array + an appended int32 property, the way both reference writers build it
    (dtype.descr + [('label','i4')], copy every column, set the label column;
    deep_learning_segmentation.py:318-328, k_means.py:181-187)."""
    descr = [(n, vertex.dtype[n].str) for n in vertex.dtype.names] + [(name, "<i4")]
    out = np.empty(len(vertex), dtype=descr)
    for n in vertex.dtype.names:
        out[n] = vertex[n]
    out[name] = labels
    return out
