#!/usr/bin/env python3
"""ROS 2 node: the reference's C++ junction detector node (ros2_ws/src/junction_point_detector/src/
fishnet_detector_ros.cpp:30-80) on the B200 engine — `dampenIntensity(img, -20, 15)` and
`find_junctions_not_rotated(img, 200, 2.0, false, 6)` become one `ofb_find_junctions` call; the `PointCloud` on
`/junction_detector/junctions` is what `lfn3_junction_node.py:108-115` subscribes to.

Needs rclpy / sensor_msgs / geometry_msgs (integration example; `JunctionDetectorNode` is tested without ROS in
tests/test_node_gpu.py).  The C++ node converts every message to rgb8 before the detector; so does this one."""
import numpy as np
import rclpy
from geometry_msgs.msg import Point32
from rclpy.node import Node
from sensor_msgs.msg import Image, PointCloud

from opticalflowcontainer_b200.node import JunctionDetectorNode


class JunctionDetector(Node):
    def __init__(self):
        super().__init__('junction_detector')
        self.core = JunctionDetectorNode(grid_area=200, grid_area_threshold=2.0, eps=6, dampen=(-20.0, 15.0))
        self.pub = self.create_publisher(PointCloud, '/junction_detector/junctions', 10)
        self.create_subscription(Image, '/camera/camera/color/image_raw', self.image_callback, 10)

    def image_callback(self, msg: Image):
        img = np.frombuffer(msg.data, dtype=np.uint8).reshape(msg.height, msg.step)[:, :msg.width * 3]
        img = img.reshape(msg.height, msg.width, 3)
        if msg.encoding == 'bgr8':                       # cv_bridge::toCvCopy(msg, RGB8)
            img = img[..., ::-1]
        stamp = msg.header.stamp.sec + msg.header.stamp.nanosec * 1e-9
        cloud = self.core.image_callback(np.ascontiguousarray(img), stamp, msg.header.frame_id)
        if cloud is None:
            self.get_logger().info('No junctions found')
            return
        out = PointCloud()
        out.header = msg.header
        out.points = [Point32(x=float(p[0]), y=float(p[1]), z=0.0) for p in cloud.points]
        self.pub.publish(out)


def main(args=None):
    rclpy.init(args=args)
    node = JunctionDetector()
    try:
        rclpy.spin(node)
    except KeyboardInterrupt:
        pass
    finally:
        node.destroy_node()
        rclpy.shutdown()


if __name__ == '__main__':
    main()
