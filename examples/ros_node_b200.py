#!/usr/bin/env python3
"""ROS 1 adapter: the reference node's message handling around the device-resident filter.

SURVEY 8(f) rank 2: where the calls of `mcmh_localization_b200.Localizer` go in a node shaped like
app/scripts/amcmh_localizer.py -- the reference's topics (/scan, /odom, /map in; /mcmh_estimated_pose, /mcmh_particles
out: node:104-110, 126) and rosparam keys (app/params/amhmcl.yaml).  There is no ROS in the build image, so
tests/test_ros_adapter.py runs these callbacks under stub rospy / message modules with synthetic messages.

    rosrun <pkg> ros_node_b200.py _localization_mode:=MHMCL _init_particles:=100000
"""
import math

import numpy as np

try:                                    # the filter itself needs none of these
    import rospy
    import tf
    from geometry_msgs.msg import PoseWithCovarianceStamped
    from nav_msgs.msg import OccupancyGrid, Odometry
    from sensor_msgs.msg import LaserScan
    from visualization_msgs.msg import Marker, MarkerArray
except ImportError:                     # pragma: no cover
    rospy = None

from mcmh_localization_b200 import Localizer

PARAM_KEYS = ("alpha1", "alpha2", "alpha3", "alpha4", "sigma_hit", "z_hit", "z_rand", "max_range", "step", "kld_epsilon", "kld_z",
              "kld_bin_size_xy", "kld_bin_size_theta", "min_particles", "alpha_slow", "alpha_fast")      # amhmcl.yaml


def yaw_of(q):
    return math.atan2(2.0 * (q.w * q.z + q.x * q.y), 1.0 - 2.0 * (q.y * q.y + q.z * q.z))


class B200LocalizerNode:
    def __init__(self):
        params = {k: rospy.get_param("~" + k) for k in PARAM_KEYS if rospy.has_param("~" + k)}
        self.n = int(rospy.get_param("~init_particles", 2000))                      # node:26
        self.max_markers = int(rospy.get_param("~max_markers", 2000))
        # node:18 localization_mode; the resampling arithmetic is the reference's own (pu:416-446, bit-exact)
        self.loc = Localizer(params=params, mode=rospy.get_param("~localization_mode", "MHAMCL"))
        self.ready = False
        self.pose_pub = rospy.Publisher("/mcmh_estimated_pose", PoseWithCovarianceStamped, queue_size=1)     # node:109
        self.marker_pub = rospy.Publisher("/mcmh_particles", MarkerArray, queue_size=1)
        rospy.Subscriber("/map", OccupancyGrid, self.on_map, queue_size=1)
        rospy.Subscriber("/odom", Odometry, self.on_odom, queue_size=10)
        rospy.Subscriber("/scan", LaserScan, self.on_scan, queue_size=1)

    def on_map(self, msg):                       # node:124-177 load_map
        info = msg.info
        occ = np.asarray(msg.data, dtype=np.int8).reshape(info.height, info.width)
        self.loc.load_map(occ, info.resolution, (info.origin.position.x, info.origin.position.y))
        self.loc.init_uniform(self.n)            # node:188 generate_valid_particles
        self.ready = True

    def on_odom(self, msg):                      # node:379-408 odom_callback / move_particles
        if self.ready:
            p = msg.pose.pose
            self.loc.predict((p.position.x, p.position.y, yaw_of(p.orientation)))

    def on_scan(self, msg):                      # node:294-338 lidar_callback
        if not self.ready:
            return
        self.loc.update(np.asarray(msg.ranges, dtype=np.float32), msg.angle_min, msg.angle_max)
        est = self.loc.estimate()                # node:586-597
        self.loc.resample()                      # node:488-492 / node:496-527 by mode
        if est is None:                          # node:594-596: fewer than two particles, nothing is published
            return
        mx, my, mt, cov = est
        out = PoseWithCovarianceStamped()
        out.header.stamp, out.header.frame_id = msg.header.stamp, "map"
        out.pose.pose.position.x, out.pose.pose.position.y = mx, my
        qz, qw = math.sin(0.5 * mt), math.cos(0.5 * mt)
        out.pose.pose.orientation.z, out.pose.pose.orientation.w = qz, qw
        c = [0.0] * 36                           # node:608-619: x, y, yaw rows / columns of the 6 x 6 covariance
        for a, i in enumerate((0, 1, 5)):
            for b, j in enumerate((0, 1, 5)):
                c[6 * i + j] = float(cov[a][b])
        out.pose.covariance = c
        self.pose_pub.publish(out)
        self.publish_markers(msg.header.stamp)

    def publish_markers(self, stamp):            # node:538-581, on a device-side sample of the cloud
        poses, w = self.loc.particles_sample(self.max_markers)
        span = float(w.max() - w.min()) + 1e-6
        arr = MarkerArray()
        clear = Marker()
        clear.action = Marker.DELETEALL
        arr.markers.append(clear)
        for k, (p, wk) in enumerate(zip(poses, w)):
            m = Marker()
            m.header.frame_id, m.header.stamp = "map", stamp
            m.ns, m.id, m.type, m.action = "particles", k, Marker.ARROW, Marker.ADD
            m.scale.x, m.scale.y, m.scale.z = 0.1, 0.02, 0.02
            s = (float(wk) - float(w.min())) / span
            m.color.a, m.color.r, m.color.g, m.color.b = 1.0, s, 0.0, 1.0 - s
            m.pose.position.x, m.pose.position.y = float(p[0]), float(p[1])
            m.pose.orientation.z, m.pose.orientation.w = math.sin(0.5 * p[2]), math.cos(0.5 * p[2])
            arr.markers.append(m)
        self.marker_pub.publish(arr)


def main():
    if rospy is None:
        raise SystemExit("this adapter needs a ROS 1 environment (rospy, tf, message packages)")
    rospy.init_node("mcmh_localizer_b200")
    B200LocalizerNode()
    rospy.spin()


if __name__ == "__main__":
    main()
