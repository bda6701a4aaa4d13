"""PCD files as the reference writes and reads them for keyframes: `pcl::io::savePCDFileBinary(path + ".pcd", *cloud)` and
`pcl::io::loadPCDFile` on pcl::PointCloud<pcl::PointXYZI> (/root/reference/src/mrg_slam/keyframe.cpp:108-110,195-197).

PCL writes the raw array of 32-byte PointXYZI structs and describes the padding as `_` fields:
    FIELDS x y z _ intensity _ / SIZE 4 4 4 1 4 1 / TYPE F F F U F U / COUNT 1 1 1 4 1 12 / DATA binary
The reader follows the header generically (any field order, float32 / float64 / integer fields, ascii or binary data), so
KITTI-style 16-byte x y z intensity files load too; binary_compressed is refused.
Returns / takes (n, 4) float32 arrays x, y, z, intensity — the layout the engine's C ABI uses.
"""
import numpy as np

_TYPES = {("F", 4): "<f4", ("F", 8): "<f8", ("U", 1): "u1", ("U", 2): "<u2", ("U", 4): "<u4", ("U", 8): "<u8",
          ("I", 1): "i1", ("I", 2): "<i2", ("I", 4): "<i4", ("I", 8): "<i8"}


def write_pcd(path, points, pcl_layout=True):
    """Writes `points` ((n,4) float32 x,y,z,intensity) the way savePCDFileBinary does for PointXYZI (32 B/point, padding
    described by `_` fields) or, with pcl_layout=False, packed 16 B/point."""
    a = np.ascontiguousarray(points, dtype=np.float32)
    n = len(a)
    if pcl_layout:
        hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z _ intensity _\nSIZE 4 4 4 1 4 1\nTYPE F F F U F U\n"
               "COUNT 1 1 1 4 1 12\n")
        raw = np.zeros((n, 8), dtype=np.float32)
        raw[:, :3] = a[:, :3]
        raw[:, 3] = 1.0  # data[3] = 1.0f, as PointXYZI's constructor leaves it
        raw[:, 4] = a[:, 3]
    else:
        hdr = "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\n"
        raw = a
    hdr += f"WIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA binary\n"
    with open(path, "wb") as f:
        f.write(hdr.encode("ascii"))
        f.write(raw.tobytes())


def read_pcd(path):
    """Reads x, y, z, intensity (0 when the file has no intensity field) as an (n, 4) float32 array."""
    with open(path, "rb") as f:
        blob = f.read()
    meta, pos = {}, 0
    while True:
        end = blob.index(b"\n", pos)
        line = blob[pos:end].decode("ascii", "replace").strip()
        pos = end + 1
        if not line or line.startswith("#"):
            continue
        key, *vals = line.split()
        meta[key.upper()] = vals
        if key.upper() == "DATA":
            break
    fields = meta["FIELDS"]
    sizes = [int(v) for v in meta["SIZE"]]
    types = meta["TYPE"]
    counts = [int(v) for v in meta.get("COUNT", ["1"] * len(fields))]
    n = int(meta["POINTS"][0]) if "POINTS" in meta else int(meta["WIDTH"][0]) * int(meta["HEIGHT"][0])
    mode = meta["DATA"][0].lower()
    want = {"x": 0, "y": 1, "z": 2, "intensity": 3}
    out = np.zeros((n, 4), dtype=np.float32)
    if mode == "binary":
        dt, off = [], 0
        names, offsets, formats = [], [], []
        for name, sz, ty, cnt in zip(fields, sizes, types, counts):
            if name in want and cnt == 1:
                names.append(name); offsets.append(off); formats.append(_TYPES[(ty, sz)])
            off += sz * cnt
        rec = np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": off})
        data = np.frombuffer(blob, dtype=rec, count=n, offset=pos)
        for name in names:
            out[:, want[name]] = data[name].astype(np.float32)
    elif mode == "ascii":
        cols, c = {}, 0
        for name, cnt in zip(fields, counts):
            if name in want and cnt == 1:
                cols[name] = c
            c += cnt
        rows = np.loadtxt(blob[pos:].decode("ascii").splitlines(), ndmin=2) if n else np.zeros((0, c))
        for name, col in cols.items():
            out[:, want[name]] = rows[:n, col].astype(np.float32)
    else:
        raise ValueError(f"PCD DATA {mode} is not supported (the reference writes DATA binary)")
    return out
