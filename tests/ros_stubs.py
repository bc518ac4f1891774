"""Minimal stand-ins for the ROS 1 Python modules examples/ros_node_b200.py imports (rospy, tf and the message
packages), so that its callbacks can be executed in an image without ROS.  Only what the adapter touches exists:
parameters, publishers that record what they were given, subscribers that record their callback, and message
classes that are plain attribute bags with the fields of the real messages."""
import sys
import types


class _Bag:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class Header(_Bag):
    def __init__(self):
        super().__init__(stamp=0.0, frame_id="")


class Point(_Bag):
    def __init__(self):
        super().__init__(x=0.0, y=0.0, z=0.0)


class Quaternion(_Bag):
    def __init__(self):
        super().__init__(x=0.0, y=0.0, z=0.0, w=1.0)


class Pose(_Bag):
    def __init__(self):
        super().__init__(position=Point(), orientation=Quaternion())


class PoseWithCovariance(_Bag):
    def __init__(self):
        super().__init__(pose=Pose(), covariance=[0.0] * 36)


class PoseWithCovarianceStamped(_Bag):
    def __init__(self):
        super().__init__(header=Header(), pose=PoseWithCovariance())


class Odometry(_Bag):
    def __init__(self):
        super().__init__(header=Header(), pose=PoseWithCovariance())


class MapMetaData(_Bag):
    def __init__(self):
        super().__init__(resolution=0.05, width=0, height=0, origin=Pose())


class OccupancyGrid(_Bag):
    def __init__(self):
        super().__init__(header=Header(), info=MapMetaData(), data=[])


class LaserScan(_Bag):
    def __init__(self):
        super().__init__(header=Header(), angle_min=0.0, angle_max=0.0, angle_increment=0.0, range_min=0.0, range_max=0.0,
                         ranges=[])


class Marker(_Bag):
    ARROW, ADD, DELETEALL = 0, 0, 3

    def __init__(self):
        super().__init__(header=Header(), ns="", id=0, type=0, action=0, pose=Pose(), scale=Point(),
                         color=_Bag(r=0.0, g=0.0, b=0.0, a=0.0))


class MarkerArray(_Bag):
    def __init__(self):
        super().__init__(markers=[])


class Publisher:
    def __init__(self, topic, cls, queue_size=1):
        self.topic, self.cls, self.sent = topic, cls, []
        PUBLISHERS[topic] = self

    def publish(self, msg):
        self.sent.append(msg)


class Subscriber:
    def __init__(self, topic, cls, callback, queue_size=1):
        self.topic, self.cls, self.callback = topic, cls, callback
        SUBSCRIBERS[topic] = self


PARAMS, PUBLISHERS, SUBSCRIBERS = {}, {}, {}


def install(params=None):
    """Put the stub modules into sys.modules (idempotent) and reset the recorded state."""
    PARAMS.clear(); PUBLISHERS.clear(); SUBSCRIBERS.clear()
    PARAMS.update(params or {})
    rospy = types.ModuleType("rospy")
    rospy.get_param = lambda name, default=None: PARAMS.get(name, default)
    rospy.has_param = lambda name: name in PARAMS
    rospy.Publisher, rospy.Subscriber = Publisher, Subscriber
    rospy.init_node = lambda name: None
    rospy.spin = lambda: None
    rospy.loginfo = rospy.logwarn = rospy.logerr = lambda *a, **k: None
    mods = {"rospy": rospy, "tf": types.ModuleType("tf")}
    for pkg, classes in (("geometry_msgs", (PoseWithCovarianceStamped,)), ("nav_msgs", (OccupancyGrid, Odometry)),
                         ("sensor_msgs", (LaserScan,)), ("visualization_msgs", (Marker, MarkerArray))):
        top, msg = types.ModuleType(pkg), types.ModuleType(pkg + ".msg")
        for c in classes:
            setattr(msg, c.__name__, c)
        top.msg = msg
        mods[pkg], mods[pkg + ".msg"] = top, msg
    sys.modules.update(mods)
    return rospy
