"""Minimal reader of the .vtu files the output path writes (UnstructuredGrid, inline base64 binary, uncompressed,
UInt32 headers): used by the tests and handy for quick looks without ParaView."""
import base64
import xml.etree.ElementTree as ET

import numpy as np

_DTYPES = {"Float64": np.float64, "Float32": np.float32, "Int32": np.int32, "Int64": np.int64, "UInt8": np.uint8}


def _decode(elem):
    raw = base64.b64decode("".join(elem.text.split()))
    n = int(np.frombuffer(raw[:4], np.uint32)[0])
    a = np.frombuffer(raw[4:4 + n], _DTYPES[elem.get("type")])
    comps = int(elem.get("NumberOfComponents", "1"))
    return a.reshape(-1, comps) if comps > 1 else a


def read_vtu(path):
    """-> dict(points [np,3], connectivity, offsets, types, point_data {name: array})"""
    root = ET.parse(path).getroot()
    if root.get("type") != "UnstructuredGrid" or root.get("byte_order") != "LittleEndian":
        raise ValueError(f"{path}: not a little-endian UnstructuredGrid file")
    piece = root.find("UnstructuredGrid/Piece")
    out = {"n_points": int(piece.get("NumberOfPoints")), "n_cells": int(piece.get("NumberOfCells")),
           "points": _decode(piece.find("Points/DataArray")), "point_data": {}}
    for a in piece.findall("Cells/DataArray"):
        out[a.get("Name")] = _decode(a)
    for a in piece.findall("PointData/DataArray"):
        out["point_data"][a.get("Name")] = _decode(a)
    return out
