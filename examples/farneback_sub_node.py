#!/usr/bin/env python3
"""ROS 2 subscriber node: the reference's `lfn3_sub_node.py` (ros2_ws/src/liteflownet3/liteflownet3/
lfn3_sub_node.py:141-222) with the neural flow call replaced by the B200 Farneback engine.

Needs rclpy / sensor_msgs / geometry_msgs (not available in the build container, so this file is an
integration example; the frame -> velocity logic it delegates to, `FarnebackVelocityNode`, is tested
without ROS in tests/test_node_gpu.py)."""
import numpy as np
import rclpy
from geometry_msgs.msg import Vector3Stamped
from rclpy.node import Node
from sensor_msgs.msg import Image

from opticalflowcontainer_b200.node import FarnebackVelocityNode


class FarnebackSubNode(Node):
    def __init__(self):
        super().__init__('farneback_sub_node')
        for name, default in (('width', 640), ('height', 480), ('pixel_to_meter', 0.0011), ('reduce', 'median'),
                              ('device', 0), ('pyr_scale', 0.5), ('levels', 3), ('winsize', 15), ('iterations', 3),
                              ('poly_n', 5), ('poly_sigma', 1.2), ('flags', 0), ('on_device_reduce', True)):
            self.declare_parameter(name, default)
        g = lambda n: self.get_parameter(n).value  # noqa: E731
        self.core = FarnebackVelocityNode(width=g('width'), height=g('height'), pixel_to_meter=g('pixel_to_meter'),
                                          reduce=g('reduce'), device=g('device'), pyr_scale=g('pyr_scale'),
                                          levels=g('levels'), winsize=g('winsize'), iterations=g('iterations'),
                                          poly_n=g('poly_n'), poly_sigma=g('poly_sigma'), flags=g('flags'),
                                          on_device_reduce=g('on_device_reduce'))
        self.flow_pub = self.create_publisher(Vector3Stamped, '/optical_flow/farneback_velocity', 10)
        self.smooth_flow_pub = self.create_publisher(Vector3Stamped, '/optical_flow/farneback_smooth_velocity', 10)
        self.create_subscription(Image, '/camera/camera/color/image_raw', self.image_callback, 10)

    def image_callback(self, msg: Image):
        try:
            img = np.frombuffer(msg.data, dtype=np.uint8).reshape(msg.height, msg.step)[:, :msg.width * 3]
            img = img.reshape(msg.height, msg.width, 3)
            stamp = msg.header.stamp.sec + msg.header.stamp.nanosec * 1e-9
            out = self.core.image_callback(img, stamp, msg.encoding)
        except Exception as e:  # same failure policy as the reference nodes: log and drop the frame
            self.get_logger().error(f"Error computing optical flow: {e}")
            return
        if out is None:
            return
        for pub, m in zip((self.flow_pub, self.smooth_flow_pub), out):
            v = Vector3Stamped()
            v.header.stamp = msg.header.stamp
            v.header.frame_id = m.frame_id
            v.vector.x, v.vector.y, v.vector.z = m.vector
            pub.publish(v)


def main(args=None):
    rclpy.init(args=args)
    node = FarnebackSubNode()
    try:
        rclpy.spin(node)
    except KeyboardInterrupt:
        pass
    finally:
        node.destroy_node()
        rclpy.shutdown()


if __name__ == '__main__':
    main()
