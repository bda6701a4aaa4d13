"""PCD keyframe files (keyframe.cpp:108-110,195-197) and the other-robot point removal
(mrg_slam_component.cpp:395-427)."""
import os
import subprocess

import numpy as np
import pytest

from mrg_slam_b200 import pcd, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pcd_roundtrip_pcl_layout(tmp_path):
    a = synth.scan(synth.VLP16, 2)[:5000]
    p = str(tmp_path / "000000.pcd")
    pcd.write_pcd(p, a)
    blob = open(p, "rb").read()
    assert b"FIELDS x y z _ intensity _" in blob and b"COUNT 1 1 1 4 1 12" in blob and b"DATA binary" in blob
    assert len(blob) == blob.index(b"DATA binary\n") + len(b"DATA binary\n") + 32 * len(a)  # raw 32-byte PointXYZI structs
    b = pcd.read_pcd(p)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_pcd_packed_and_ascii(tmp_path):
    a = synth.scan(synth.VLP16, 2)[:300]
    p = str(tmp_path / "packed.pcd")
    pcd.write_pcd(p, a, pcl_layout=False)
    assert np.array_equal(pcd.read_pcd(p), a)
    q = str(tmp_path / "ascii.pcd")
    with open(q, "w") as f:
        f.write("# .PCD v0.7\nVERSION 0.7\nFIELDS intensity x y z\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\nWIDTH 3\nHEIGHT 1\n"
                "VIEWPOINT 0 0 0 1 0 0 0\nPOINTS 3\nDATA ascii\n0.5 1 2 3\n0.25 -1 -2 -3\n0 4.5 5.5 6.5\n")
    assert np.array_equal(pcd.read_pcd(q), np.array([[1, 2, 3, 0.5], [-1, -2, -3, 0.25], [4.5, 5.5, 6.5, 0]], dtype=np.float32))
    z = str(tmp_path / "xyz.pcd")  # no intensity field: zeros
    with open(z, "w") as f:
        f.write("VERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA ascii\n1 2 3\n")
    assert np.array_equal(pcd.read_pcd(z), np.array([[1, 2, 3, 0]], dtype=np.float32))
    bad = str(tmp_path / "c.pcd")
    with open(bad, "w") as f:
        f.write("VERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA binary_compressed\n")
    with pytest.raises(ValueError):
        pcd.read_pcd(bad)


def test_cpp_pcd_io_matches_python(tmp_path):
    """The C++ header writes the same bytes and reads the Python-written file."""
    src = tmp_path / "t.cpp"
    src.write_text('''
#include "b2r/pcd_io.hpp"
int main(int argc, char** argv) {
  b2r::PointCloud c;
  if (!b2r::load_pcd(argv[1], c)) return 2;
  if (!b2r::save_pcd_binary(argv[2], c)) return 3;
  std::printf("%zu\\n", c.size());
  return 0;
}
''')
    exe = tmp_path / "t"
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-ldl"], check=True)
    a = synth.scan(synth.VLP16, 9)[:4000]
    p1, p2 = str(tmp_path / "a.pcd"), str(tmp_path / "b.pcd")
    pcd.write_pcd(p1, a)
    r = subprocess.run([str(exe), p1, p2], check=True, capture_output=True, text=True)
    assert int(r.stdout) == len(a)
    assert open(p1, "rb").read() == open(p2, "rb").read()


def oracle_remove(cloud, others_map, map2sensor, radius):
    """mrg_slam_component.cpp:395-427 restated in numpy float32 (squaredNorm = x^2 + (y^2 + z^2))."""
    sensor = (np.asarray(others_map, np.float64) @ map2sensor[:3, :3].T + map2sensor[:3, 3]).astype(np.float32)
    r2 = np.float32(np.float64(radius) * np.float64(radius))
    hit = np.zeros(len(cloud), dtype=bool)
    for o in sensor:
        d = cloud[:, :3] - o
        sq = d[:, 0] * d[:, 0] + (d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])
        hit |= sq < r2
    return cloud[~hit], cloud[hit]


@pytest.mark.gpu
def test_gpu_robot_point_removal_bit_exact():
    from mrg_slam_b200 import lib as B
    reg = B.Registration(B.default_config(B.FAST_GICP))
    cloud = synth.scan(synth.HDL64, 12)
    M = np.eye(4); M[:3, 3] = [0.3, -0.2, 0.05]
    c, s = np.cos(0.3), np.sin(0.3)
    M[:3, :3] = [[c, -s, 0], [s, c, 0], [0, 0, 1]]
    others = np.array([[6.0, 1.0, -1.0], [-4.0, 3.5, -1.2], [100.0, 0.0, 0.0]])
    for radius in (2.0, 0.5):
        kept, removed = reg.remove_robot_points(cloud, others, M, radius)
        ok, orr = oracle_remove(cloud, others, M, radius)
        assert len(removed) > 0 or radius < 1
        assert np.array_equal(kept.view(np.uint32), ok.view(np.uint32)) and np.array_equal(removed.view(np.uint32), orr.view(np.uint32))
    kept, removed = reg.remove_robot_points(cloud, np.zeros((0, 3)), M, 2.0)  # no other robots: untouched (:394)
    assert np.array_equal(kept, cloud) and len(removed) == 0
