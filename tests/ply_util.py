"""Test helper: writes binary little-endian PLY files with an arbitrary vertex property mix."""
import numpy as np

NP_TYPES = {"char": "i1", "uchar": "u1", "short": "<i2", "ushort": "<u2", "int": "<i4", "uint": "<u4", "float": "<f4", "double": "<f8",
            "int8": "i1", "uint8": "u1", "int16": "<i2", "uint16": "<u2", "int32": "<i4", "uint32": "<u4", "float32": "<f4", "float64": "<f8"}


def random_column(rng, typ, n, coord=False):
    dt = np.dtype(NP_TYPES[typ])
    if dt.kind == "f":
        v = (rng.random(n) * 2000 - 1000) if coord else rng.random(n)
        return v.astype(dt)
    info = np.iinfo(dt)
    return rng.integers(info.min, info.max, n, dtype=dt, endpoint=True)


def write_ply(path, n, props, rng, offset=None, comments=(), trailing_element=True, columns=None, fmt="binary_little_endian", body_cut=0):
    """props: list of (type, name).  Returns {name: column}.  `columns` overrides generated data."""
    cols = {}
    fields = []
    for typ, name in props:
        c = columns[name] if columns and name in columns else random_column(rng, typ, n, coord=name in "xyz")
        cols[name] = c
        fields.append((name, NP_TYPES[typ]))
    rec = np.zeros(n, dtype=np.dtype(fields))  # packed (align=False)
    for name, _ in fields:
        rec[name] = cols[name]
    head = ["ply", "format %s 1.0" % fmt]
    head += ["comment %s" % c for c in comments]
    if offset is not None:
        head.append("comment offset: %r %r %r" % tuple(float(o) for o in offset))
    head.append("element vertex %d" % n)
    head += ["property %s %s" % (t, nm) for t, nm in props]
    if trailing_element:
        head += ["element face 0", "property list uchar int vertex_indices"]
    head.append("end_header")
    body = rec.tobytes()
    if body_cut:
        body = body[:-body_cut]
    with open(path, "wb") as f:
        f.write(("\n".join(head) + "\n").encode())
        f.write(body)
    return cols
