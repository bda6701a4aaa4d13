"""MapCloudGenerator::generate + pcl::ApproximateMeanVoxelGrid (both in the reference tree:
src/mrg_slam/map_cloud_generator.cpp:14-86, include/pcl/filters/ApproximateMeanVoxelGrid.hpp:63-126)."""
import numpy as np
import pytest

from mrg_slam_b200 import synth
from tests import oraclelib as O
from tests.conftest import oracle_prefilter


def keyframes(count, first=40, step=3):
    clouds = [oracle_prefilter(synth.scan(synth.VLP16, first + step * i)) for i in range(count)]
    p0 = np.linalg.inv(synth.pose(first))
    poses = [p0 @ synth.pose(first + step * i) for i in range(count)]
    return clouds, poses


def sort_by_key(points, keys):
    order = np.lexsort((keys[:, 0], keys[:, 1], keys[:, 2]))  # ascending z, then y, then x
    return points[order], keys[order]


def voxel_keys(points, resolution):
    inv = np.float32(1.0) / np.float32(resolution)
    return np.floor(points[:, :3] * inv).astype(np.int32)


# ---- oracle against an independent numpy restatement (float32 sums in input order via np.add.at)
def test_oracle_matches_numpy_restatement():
    clouds, poses = keyframes(3)
    res = np.float32(0.25)
    out, keys = O.map_cloud(clouds, poses, resolution=float(res), min_points_per_voxel=2, distance_far_thresh=20.0)
    world = []
    for c, P in zip(clouds, poses):
        Pf = P.astype(np.float32)
        sq = c[:, 0] * c[:, 0] + (c[:, 1] * c[:, 1] + c[:, 2] * c[:, 2])
        c = c[~(sq > np.float32(20.0) * np.float32(20.0))]
        xyz = ((Pf[:3, 0] * c[:, 0:1] + Pf[:3, 1] * c[:, 1:2]) + Pf[:3, 2] * c[:, 2:3]) + Pf[:3, 3]
        world.append(np.concatenate([xyz.astype(np.float32), c[:, 3:]], axis=1))
    world = np.concatenate(world)
    k = voxel_keys(world, res)
    uniq, inv = np.unique(k, axis=0, return_inverse=True)
    sums = np.zeros((len(uniq), 4), dtype=np.float32)
    cnt = np.zeros(len(uniq), dtype=np.int32)
    for i in range(len(world)):  # sequential float32 accumulation, as the hash map does
        sums[inv[i]] += world[i]
        cnt[inv[i]] += 1
    sel = cnt >= 2
    ref = sums[sel] / cnt[sel, None].astype(np.float32)
    ref_keys = uniq[sel]
    a, ka = sort_by_key(out, keys)
    b, kb = sort_by_key(ref, ref_keys)
    assert np.array_equal(ka, kb)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_oracle_null_and_full_resolution_cases():
    clouds, poses = keyframes(2)
    assert O.map_cloud([], [], resolution=0.1) is None                       # no keyframes (:20-23)
    out, _ = O.map_cloud(clouds, poses, resolution=0.0)                       # full resolution (:67-71)
    assert len(out) == len(clouds[0]) + len(clouds[1])
    assert np.array_equal(out[: len(clouds[0]), 3], clouds[0][:, 3])          # intensity carried over (:45)
    out, _ = O.map_cloud(clouds, poses, first_keyframe=[1, 0], resolution=0.0, skip_first_cloud=True)
    assert len(out) == len(clouds[1])                                         # first keyframe skipped (:33-35)
    assert O.map_cloud(clouds, poses, resolution=0.1, distance_far_thresh=1e-3) is None   # everything filtered, >1 keyframes (:58-61)


@pytest.mark.gpu
@pytest.mark.parametrize("resolution,min_pts,far", [(0.05, 1, -1.0), (0.25, 2, 20.0), (1.0, 3, 35.0)])
def test_gpu_map_cloud_bit_exact(resolution, min_pts, far):
    from mrg_slam_b200 import lib as B
    clouds, poses = keyframes(5)
    first = [1, 0, 0, 0, 0]
    reg = B.Registration(B.default_config(B.FAST_VGICP))
    for skip in (False, True):
        g = reg.map_cloud(clouds, poses, first, resolution, min_pts, far, skip)
        o, ko = O.map_cloud(clouds, poses, first, resolution, min_pts, far, skip)
        o, ko = sort_by_key(o, ko)
        assert len(g) == len(o)
        assert np.array_equal(g.view(np.uint32), o.view(np.uint32))          # same voxel set, bit-identical means, (z, y, x) order


@pytest.mark.gpu
def test_gpu_map_cloud_edge_cases():
    from mrg_slam_b200 import lib as B
    clouds, poses = keyframes(2)
    reg = B.Registration(B.default_config(B.FAST_VGICP))
    assert reg.map_cloud([], [], resolution=0.1) is None
    full = reg.map_cloud(clouds, poses, resolution=0.0)
    ofull, _ = O.map_cloud(clouds, poses, resolution=0.0)
    assert np.array_equal(full.view(np.uint32), ofull.view(np.uint32))        # unfiltered cloud: same points, same order
    assert reg.map_cloud(clouds, poses, resolution=0.1, distance_far_thresh=1e-3) is None
    one = reg.map_cloud(clouds[:1], poses[:1], resolution=0.1, distance_far_thresh=1e-3)
    assert one is not None and len(one) == 0                                  # a single keyframe yields an empty cloud, not nullptr
    # 32-byte pcl::PointXYZI stride
    c32 = np.zeros((len(clouds[0]), 8), dtype=np.float32)
    c32[:, :3] = clouds[0][:, :3]; c32[:, 3] = 1.0; c32[:, 4] = clouds[0][:, 3]
    a = reg.map_cloud([c32], poses[:1], resolution=0.2)
    b = reg.map_cloud(clouds[:1], poses[:1], resolution=0.2)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
