"""examples/ros_node_b200.py (the ROS adapter around the device-resident filter) executed under
stub rospy / message modules (tests/ros_stubs.py) with synthetic /map, /odom and /scan messages: the callbacks run
on the GPU exactly as they would under ROS; only the transport is faked."""
import importlib.util
import math
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu


def _load_adapter(params):
    import ros_stubs
    ros_stubs.install(params)
    spec = importlib.util.spec_from_file_location("ros_node_b200", os.path.join(ROOT, "examples", "ros_node_b200.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.rospy is not None
    return mod, ros_stubs


@pytest.mark.parametrize("mode", ["MHMCL", "MCL", "AMHAMCL"])
def test_ros_adapter_callbacks_run_with_synthetic_messages(mode):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from mcmh_localization_b200.maps import load_npz
    from mcmh_localization_b200.synth import raycast_scan
    gm = load_npz(os.path.join(GOLDEN, "map_world.npz"))
    mod, stubs = _load_adapter({"~localization_mode": mode, "~init_particles": 20000, "~max_markers": 300,
                                "~sigma_hit": 0.3, "~z_hit": 0.75, "~z_rand": 0.25, "~max_range": 5.0, "~step": 1,
                                "~alpha1": 0.002, "~alpha2": 0.03, "~alpha3": 0.08, "~alpha4": 0.002})
    node = mod.B200LocalizerNode()
    assert set(stubs.SUBSCRIBERS) == {"/map", "/odom", "/scan"}                  # node:104-105, 126
    assert set(stubs.PUBLISHERS) == {"/mcmh_estimated_pose", "/mcmh_particles"}   # node:109-110
    # before the map arrives the callbacks are no-ops
    stubs.SUBSCRIBERS["/odom"].callback(stubs.Odometry())
    stubs.SUBSCRIBERS["/scan"].callback(stubs.LaserScan())
    assert not stubs.PUBLISHERS["/mcmh_estimated_pose"].sent
    grid = stubs.OccupancyGrid()
    grid.info.resolution, grid.info.width, grid.info.height = gm.resolution, gm.width, gm.height
    grid.info.origin.position.x, grid.info.origin.position.y = gm.origin_x, gm.origin_y
    grid.data = gm.occ.ravel().tolist()
    stubs.SUBSCRIBERS["/map"].callback(grid)
    assert node.ready and node.loc.n == 20000
    pose = np.array([-2.0, -0.5, 0.0])
    for k in range(8):
        od = stubs.Odometry()
        od.pose.pose.position.x, od.pose.pose.position.y = float(pose[0]), float(pose[1])
        od.pose.pose.orientation.z, od.pose.pose.orientation.w = math.sin(0.5 * pose[2]), math.cos(0.5 * pose[2])
        stubs.SUBSCRIBERS["/odom"].callback(od)
        ranges, angles = raycast_scan(gm, pose, noise_sigma=0.01, seed=k)
        sc = stubs.LaserScan()
        sc.header.stamp = 0.2 * k
        sc.angle_min, sc.angle_max = float(angles[0]), float(angles[-1])
        sc.ranges = ranges.tolist()
        stubs.SUBSCRIBERS["/scan"].callback(sc)
        pose = pose + np.array([0.02 * np.cos(pose[2]), 0.02 * np.sin(pose[2]), 0.01])
    sent = stubs.PUBLISHERS["/mcmh_estimated_pose"].sent
    assert len(sent) == 8
    for m in sent:
        assert m.header.frame_id == "map"
        x, y = m.pose.pose.position.x, m.pose.pose.position.y
        assert gm.origin_x <= x <= gm.origin_x + gm.width * gm.resolution and gm.origin_y <= y <= gm.origin_y + gm.height * gm.resolution
        c = np.array(m.pose.covariance).reshape(6, 6)
        cov = c[np.ix_((0, 1, 5), (0, 1, 5))]
        assert np.all(np.isfinite(cov)) and np.allclose(cov, cov.T, atol=1e-9) and np.all(np.linalg.eigvalsh(cov) > -1e-9)
        assert abs(m.pose.pose.orientation.z ** 2 + m.pose.pose.orientation.w ** 2 - 1.0) < 1e-12
    arrays = stubs.PUBLISHERS["/mcmh_particles"].sent
    assert len(arrays) == 8 and all(1 < len(a.markers) <= 301 for a in arrays)
    assert arrays[-1].markers[0].action == stubs.Marker.DELETEALL
