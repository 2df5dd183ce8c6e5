"""Byte-level known answers for the two labelled-PLY writers.

The reference writes through `plyfile` (deep_learning_segmentation.py:311-332 binary,
3D_clustering/k_means.py:169-194 ASCII), which is not installed here.  The contract that matters
is the consumer's: the WebGL viewer parses the binary file itself
(Web_Viewer_Gaussians_Selection/gaussians_selection.js:464-511: header up to "end_header\\n",
`element vertex N`, one `property <type> <name>` per line, packed little-endian rows; :579 reads
`label` with getInt32).  `viewer_parse` below restates that parser; the ASCII file is compared with
literal text."""
import io
import re
import struct

import numpy as np

from util import pkg

# gaussians_selection.js:487-495: type name -> DataView getter (size in bytes, struct code); anything else is getInt8
TYPE_MAP = {"double": (8, "d"), "int": (4, "i"), "uint": (4, "I"), "float": (4, "f"),
            "short": (2, "h"), "ushort": (2, "H"), "uchar": (1, "B")}


def viewer_parse(buf: bytes):
    """processPlyBuffer (gaussians_selection.js:464-511), field access as in the `attrs` proxy (:506-511)."""
    header = buf[:10 * 1024].decode("utf-8", "replace")
    end = header.index("end_header\n")                                  # :469, byte index == char index for ASCII headers
    count = int(re.search(r"element vertex (\d+)\n", header).group(1))  # :476
    offsets, codes, row = {}, {}, 0
    for line in header[:end].split("\n"):                               # :497-505
        if not line.startswith("property "):
            continue
        _, typ, name = line.split(" ")
        size, code = TYPE_MAP.get(typ, (1, "b"))
        offsets[name], codes[name] = row, code
        row += size
    body = buf[end + len("end_header\n"):]

    def attr(i, name):
        return struct.unpack_from("<" + codes[name], body, i * row + offsets[name])[0]
    return count, row, offsets, attr


def _vertices():
    v = np.zeros(3, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("opacity", "<f4"), ("scale_0", "<f4")])
    v["x"] = [0.1, -2.5, 1.0]
    v["y"] = [3.0, 1e-8, 16777216.0]
    v["z"] = [-0.0, 123456.789, 0.333333343267440796]
    v["opacity"] = [0.5, -1.0, 2.0]
    v["scale_0"] = [1.0, 2.0, 3.0]
    return v


def test_binary_labelled_ply_is_what_the_viewer_reads(tmp_path):
    dls, plyio = pkg("deep_learning_segmentation"), pkg("plyio")
    v = _vertices()
    labels = np.array([7, -1, 149], np.int32)
    src = plyio.PlyFile([plyio.PlyElementData("vertex", v)])
    out = tmp_path / "labelled.ply"
    dls.save_labeled_ply(str(out), src, labels)
    buf = open(out, "rb").read()
    # the header, byte for byte (plyfile's spelling of the types, properties in input order, label last)
    want_header = (b"ply\nformat binary_little_endian 1.0\nelement vertex 3\n"
                   b"property float x\nproperty float y\nproperty float z\nproperty float opacity\nproperty float scale_0\n"
                   b"property int label\nend_header\n")
    assert buf.startswith(want_header)
    assert len(buf) == len(want_header) + 3 * 24                        # packed rows, no padding
    # the body, byte for byte
    want_body = b"".join(struct.pack("<5fi", *[float(v[n][i]) for n in v.dtype.names], int(labels[i])) for i in range(3))
    assert buf[len(want_header):] == want_body
    # and through the viewer's own parsing logic
    count, row, offsets, attr = viewer_parse(buf)
    assert (count, row) == (3, 24) and offsets["label"] == 20
    assert [attr(i, "label") for i in range(3)] == [7, -1, 149]       # getInt32, little endian (:579)
    for n in v.dtype.names:
        assert [np.float32(attr(i, n)) for i in range(3)] == list(v[n])


def test_ascii_labelled_ply_known_answer(tmp_path):
    km, plyio = pkg("k_means"), pkg("plyio")
    v = _vertices()
    src = plyio.PlyFile([plyio.PlyElementData("vertex", v)])
    out = tmp_path / "clustered.ply"
    km.add_label_proberty(src, str(out), np.array([2, 0, 9], np.int64))
    text = open(out, "rb").read()
    # plyfile prints every field with '%.18g' of its float64 value (a float32 widened exactly), one
    # space between fields, one vertex per line
    want = (b"ply\nformat ascii 1.0\nelement vertex 3\n"
            b"property float x\nproperty float y\nproperty float z\nproperty float opacity\nproperty float scale_0\n"
            b"property int label\nend_header\n"
            b"0.100000001490116119 3 -0 0.5 1 2\n"
            b"-2.5 9.99999993922529029e-09 123456.7890625 -1 2 0\n"
            b"1 16777216 0.333333343267440796 2 3 9\n")
    assert text == want
    back = plyio.read_ply(str(out))
    assert back.text and np.array_equal(back["vertex"]["label"], [2, 0, 9])
    for n in v.dtype.names:
        assert np.array_equal(back["vertex"][n], v[n])                  # '%.18g' round-trips float32 exactly


def test_ascii_reader_is_vectorised_and_strict():
    plyio = pkg("plyio")
    n = 5000
    rng = np.random.default_rng(3)
    v = np.zeros(n, dtype=[("x", "<f4"), ("f_dc_0", "<f4"), ("label", "<i4")])
    v["x"], v["f_dc_0"], v["label"] = rng.standard_normal(n), rng.standard_normal(n) * 1e-5, rng.integers(-1, 150, n)
    f = io.BytesIO()
    plyio.write_ply(f, [("vertex", v)], text=True)
    f.seek(0)
    back = plyio.read_ply(f)
    assert back.text and back["vertex"].data.dtype == v.dtype and np.array_equal(back["vertex"].data, v)
    bad = f.getvalue().replace(b"element vertex 5000", b"element vertex 5001")
    try:
        plyio.read_ply(io.BytesIO(bad))
    except ValueError:
        pass
    else:
        raise AssertionError("a truncated ASCII element must be rejected")
