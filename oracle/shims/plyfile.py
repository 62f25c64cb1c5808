"""Import shim for plyfile (absent here, no network; the reference pins plyfile 0.7.2, environment.yml).
Implements the two calls fusion.py:300-318 makes -- ``PlyElement.describe(structured_array, name)`` and
``PlyData([el]).write(path)`` -- in plyfile's default output format: binary_little_endian 1.0, one ``element`` block
per PlyElement, one ``property <type> <name>`` line per field of the structured array (numpy f4 -> float, u1 -> uchar,
...), followed by the packed records.  Test infrastructure only."""
import numpy as np

_TYPES = {"i1": "char", "u1": "uchar", "i2": "short", "u2": "ushort", "i4": "int", "u4": "uint", "f4": "float",
          "f8": "double"}


class PlyElement:
    def __init__(self, name, data):
        self.name, self.data = name, data

    @staticmethod
    def describe(data, name):
        return PlyElement(name, np.asarray(data))


class PlyData:
    def __init__(self, elements, text=False, byte_order="<"):
        self.elements = list(elements)

    def write(self, path):
        with open(path, "wb") as f:
            hdr = ["ply", "format binary_little_endian 1.0"]
            for el in self.elements:
                hdr.append(f"element {el.name} {len(el.data)}")
                for field in el.data.dtype.names:
                    dt = el.data.dtype.fields[field][0]
                    hdr.append(f"property {_TYPES[dt.str[1:]]} {field}")
            hdr.append("end_header")
            f.write(("\n".join(hdr) + "\n").encode("ascii"))
            for el in self.elements:
                packed = np.empty(len(el.data), dtype=[(n, el.data.dtype.fields[n][0].newbyteorder("<"))
                                                        for n in el.data.dtype.names])
                for n in el.data.dtype.names:
                    packed[n] = el.data[n]
                f.write(packed.tobytes())
