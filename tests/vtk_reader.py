"""Minimal reader for the raw-appended .vtu pieces the engine writes (tests only)."""
import re

import numpy as np

_DT = {"Float64": "<f8", "Int32": "<i4", "UInt8": "u1", "Int64": "<i8"}


def read_vtu_cell_data(path):
    blob = open(path, "rb").read()
    marker = blob.index(b'<AppendedData encoding="raw">')
    start = blob.index(b"_", marker) + 1
    header = blob[:marker].decode()
    out = {}
    for m in re.finditer(r'<DataArray type="(\w+)" Name="(\w+)"[^>]*offset="(\d+)"', header):
        typ, name, off = m.group(1), m.group(2), int(m.group(3))
        nbytes = int(np.frombuffer(blob, dtype="<u8", count=1, offset=start + off)[0])
        out[name] = np.frombuffer(blob, dtype=_DT[typ], count=nbytes // np.dtype(_DT[typ]).itemsize,
                                  offset=start + off + 8).copy()
    return out
