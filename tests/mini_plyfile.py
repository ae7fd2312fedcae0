"""TEST INFRASTRUCTURE — a minimal stand-in for the third-party `plyfile` package (not in this image), just the
surface the reference's GaussianModel.save_ply / load_ply use (gaussiansplatting/scene/gaussian_model.py:410-445,
:455-551): PlyElement.describe(structured_array, name), PlyData([el]).write(path), PlyData.read(path),
.elements[0][property] and .elements[0].properties[i].name. Written from the PLY format description (header lines
`format`, `element <name> <count>`, `property <type> <name>`; binary records in header order) and plyfile's
documented defaults (binary, native byte order -> `binary_little_endian` here, no comments). With it the reference's own
save / load code runs unchanged in tests/test_ply.py against dge_b200/ply.py."""
import numpy as np

_TYPES = {"char": "i1", "uchar": "u1", "short": "i2", "ushort": "u2", "int": "i4", "uint": "u4", "float": "f4",
          "double": "f8", "int8": "i1", "uint8": "u1", "int16": "i2", "uint16": "u2", "int32": "i4", "uint32": "u4",
          "float32": "f4", "float64": "f8"}
_NAMES = {"i1": "char", "u1": "uchar", "i2": "short", "u2": "ushort", "i4": "int", "u4": "uint", "f4": "float", "f8": "double"}


class PlyProperty:
    def __init__(self, name, dtype):
        self.name, self.dtype = name, dtype


class PlyElement:
    def __init__(self, name, data):
        self.name, self.data = name, data
        self.properties = [PlyProperty(n, data.dtype[n].str.lstrip("<>=|")) for n in data.dtype.names]

    @staticmethod
    def describe(data, name):
        return PlyElement(name, np.asarray(data))

    def __getitem__(self, key):
        return self.data[key]

    @property
    def count(self):
        return len(self.data)


class PlyData:
    def __init__(self, elements, text=False, byte_order="<"):
        self.elements, self.text = list(elements), text

    def write(self, path):
        with open(path, "wb") as fh:
            head = ["ply", "format binary_little_endian 1.0"]
            for el in self.elements:
                head.append(f"element {el.name} {el.count}")
                head += [f"property {_NAMES[p.dtype]} {p.name}" for p in el.properties]
            head.append("end_header")
            fh.write(("\n".join(head) + "\n").encode("ascii"))
            for el in self.elements:
                fh.write(el.data.astype(el.data.dtype.newbyteorder("<")).tobytes())

    @staticmethod
    def read(path):
        raw = open(path, "rb").read()
        end = raw.index(b"end_header\n") + len(b"end_header\n")
        lines = [ln.strip() for ln in raw[:end].decode("ascii").splitlines()]
        assert lines[0] == "ply"
        fmt = [ln for ln in lines if ln.startswith("format")][0].split()[1]
        assert fmt in ("binary_little_endian", "binary_big_endian"), "only binary files are needed here"
        order = "<" if fmt == "binary_little_endian" else ">"
        elements, cur = [], None
        for ln in lines:
            tok = ln.split()
            if not tok or tok[0] in ("ply", "format", "comment", "obj_info", "end_header"):
                continue
            if tok[0] == "element":
                cur = [tok[1], int(tok[2]), []]
                elements.append(cur)
            elif tok[0] == "property":
                assert tok[1] != "list", "list properties are not needed here"
                cur[2].append((tok[2], order + _TYPES[tok[1]]))
        out, off = [], end
        for name, count, fields in elements:
            dt = np.dtype(fields)
            out.append(PlyElement(name, np.frombuffer(raw, dt, count=count, offset=off)))
            off += count * dt.itemsize
        assert off == len(raw), "trailing bytes"
        return PlyData(out)
