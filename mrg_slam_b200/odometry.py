"""Host mirror of ScanMatchingOdometryComponent::matching() — the caller of the registration hot path
(/root/reference/apps/scan_matching_odometry_component.cpp:195-350), the serial odometry loop BASELINE configs[1] runs.

The keyframe / initial-guess state machine is restated over anything that offers the pcl::Registration surface
(`setInputTarget`, `setInputSource`, `align`, `hasConverged`, `getFinalTransformation`): `mrg_slam_b200.lib.Registration`
(the product) and the CPU oracle used by the tests drive the very same code, so their trajectories can be compared
scan by scan.  Matrices are float32 4x4 like the reference's Eigen::Matrix4f state (`prev_trans_`, `keyframe_pose_`).

Not restated (out of scope): the IMU / robot-odometry initial guesses (`msf_delta` stays identity,
:214-260), tf broadcasting, status publication.
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class OdometryParams:
    """Defaults of config/mrg_slam.yaml:80-93 (scan_matching_odometry_component section)."""
    keyframe_delta_translation: float = 1.0   # :80
    keyframe_delta_angle: float = 0.5236      # :81
    keyframe_delta_time: float = 10000.0      # :82
    enable_transform_thresholding: bool = False
    max_acceptable_translation: float = 1.0
    max_acceptable_angle: float = 1.0
    max_consecutive_rejections: int = 5  # scan_matching_odometry_component.cpp:110, config/mrg_slam.yaml


def _quat_w(R):
    """w of Eigen::Quaternionf(R) (Shepperd's branches, float32 like the reference)."""
    R = np.asarray(R, dtype=np.float32)
    t = np.float32(R[0, 0] + R[1, 1] + R[2, 2])
    if t > 0:
        return np.float32(0.5) * np.sqrt(t + np.float32(1.0), dtype=np.float32)
    i = int(np.argmax([R[0, 0], R[1, 1], R[2, 2]]))
    j, k = (i + 1) % 3, (i + 2) % 3
    s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + np.float32(1.0), dtype=np.float32)
    return np.float32((R[k, j] - R[j, k]) * (np.float32(0.5) / s))


def rotation_angle_half(R):
    """`std::acos(Eigen::Quaternionf(R).w())` as the reference computes delta_angle (:282,:318) — half the rotation angle."""
    return float(np.arccos(np.clip(np.float64(_quat_w(R)), -1.0, 1.0)))


class ScanMatchingOdometry:
    """matching() state machine; `downsample` mirrors the component's own downsample() hook (:168-186, default NONE here
    because the prefiltering component already voxelised the cloud)."""

    def __init__(self, registration, params=None, make_cloud=None, downsample=None):
        self.reg = registration
        self.p = params or OdometryParams()
        self.make_cloud = make_cloud or (lambda pts: pts)   # e.g. lambda pts: lib.Cloud(reg, pts) keeps structures on the device
        self.downsample = downsample or (lambda pts: pts)
        self.keyframe_cloud = None
        self.keyframe_pose = np.eye(4, dtype=np.float32)
        self.keyframe_stamp = 0.0
        self.prev_trans = np.eye(4, dtype=np.float32)
        self.prev_time = None
        self.consecutive_rejections = 0
        self.keyframe_switches = 0
        self.not_converged = 0

    def matching(self, stamp, cloud):
        """Returns the odometry pose (4x4 float32, odom <- sensor) of `cloud`."""
        if self.keyframe_cloud is None:                                     # :197-205
            self.prev_time = None
            self.prev_trans = np.eye(4, dtype=np.float32)
            self.keyframe_pose = np.eye(4, dtype=np.float32)
            self.keyframe_stamp = stamp
            self.keyframe_cloud = self.make_cloud(self.downsample(cloud))
            self.reg.setInputTarget(self.keyframe_cloud)
            return np.eye(4, dtype=np.float32)

        filtered = self.make_cloud(self.downsample(cloud))                  # :207-208
        self.reg.setInputSource(filtered)
        self.reg.align(self.prev_trans)                                     # :265-266 (msf_delta = identity)

        if not self.reg.hasConverged():                                     # :270-273
            self.not_converged += 1
            return (self.keyframe_pose @ self.prev_trans).astype(np.float32)

        trans = np.asarray(self.reg.getFinalTransformation(), dtype=np.float32)   # :275-276
        odom = (self.keyframe_pose @ trans).astype(np.float32)

        if self.p.enable_transform_thresholding:                            # :278-313
            delta = (np.linalg.inv(self.prev_trans.astype(np.float64)) @ trans).astype(np.float32)
            dx = float(np.linalg.norm(delta[:3, 3]))
            da = rotation_angle_half(delta[:3, :3])
            if dx > self.p.max_acceptable_translation or da > self.p.max_acceptable_angle:
                self.consecutive_rejections += 1
                if self.consecutive_rejections >= self.p.max_consecutive_rejections:
                    self._switch_keyframe(filtered, odom, stamp)
                    self.consecutive_rejections = 0
                    return self.keyframe_pose
                self.prev_time = stamp
                return (self.keyframe_pose @ self.prev_trans).astype(np.float32)
            self.consecutive_rejections = 0

        self.prev_time = stamp                                              # :315-316
        self.prev_trans = trans

        delta_translation = float(np.linalg.norm(trans[:3, 3]))            # :324-335
        delta_angle = rotation_angle_half(trans[:3, :3])
        delta_time = stamp - self.keyframe_stamp
        if (delta_translation > self.p.keyframe_delta_translation or delta_angle > self.p.keyframe_delta_angle
                or delta_time > self.p.keyframe_delta_time):
            self._switch_keyframe(filtered, odom, stamp)
        return odom

    def _switch_keyframe(self, filtered, odom, stamp):
        self.keyframe_cloud = filtered
        self.reg.setInputTarget(self.keyframe_cloud)    # the source just aligned becomes the target: its structures are reused
        self.keyframe_pose = odom
        self.keyframe_stamp = stamp
        self.prev_time = stamp
        self.prev_trans = np.eye(4, dtype=np.float32)
        self.keyframe_switches += 1
