#!/usr/bin/env python3
"""ROS 2 subscriber node for `sensor_msgs/CompressedImage` (JPEG) frames: the reference's
`opticalflow_comprerssed_node.py` (ros2_ws/src/optical_flow/optical_flow/opticalflow_comprerssed_node.py:41-128) with
`cv2.imdecode` + the neural flow call replaced by the B200 engine — the JPEG bytes go to the device as they arrive
(Huffman decoding, IDCT, colour conversion, resize, gray conversion there), the flow's mean / median comes back.

Needs rclpy / sensor_msgs / geometry_msgs (integration example; `FarnebackVelocityNode.compressed_callback` is tested
without ROS in tests/test_node_gpu.py)."""
import rclpy
from geometry_msgs.msg import Vector3Stamped
from rclpy.node import Node
from sensor_msgs.msg import CompressedImage

from opticalflowcontainer_b200.node import FarnebackVelocityNode


class FarnebackCompressedNode(Node):
    def __init__(self):
        super().__init__('farneback_compressed_node')
        for name, default in (('width', 640), ('height', 480), ('pixel_to_meter', 0.000566), ('reduce', 'mean'), ('device', 0)):
            self.declare_parameter(name, default)
        g = lambda n: self.get_parameter(n).value  # noqa: E731
        self.core = FarnebackVelocityNode(width=g('width'), height=g('height'), pixel_to_meter=g('pixel_to_meter'),
                                          reduce=g('reduce'), device=g('device'), on_device_reduce=True)
        self.flow_pub = self.create_publisher(Vector3Stamped, '/optical_flow/farneback_velocity', 10)
        self.smooth_flow_pub = self.create_publisher(Vector3Stamped, '/optical_flow/farneback_smooth_velocity', 10)
        self.create_subscription(CompressedImage, '/camera/camera/color/image_raw/compressed', self.image_callback, 10)

    def image_callback(self, msg: CompressedImage):
        stamp = msg.header.stamp.sec + msg.header.stamp.nanosec * 1e-9
        try:
            out = self.core.compressed_callback(bytes(msg.data), stamp)
        except Exception as e:
            self.get_logger().error(f"Error computing optical flow: {e}")
            return
        if out is None:             # priming frame, or a stream the device path does not decode (logged by the reference too)
            return
        for pub, m in zip((self.flow_pub, self.smooth_flow_pub), out):
            v = Vector3Stamped()
            v.header.stamp = msg.header.stamp
            v.header.frame_id = m.frame_id
            v.vector.x, v.vector.y, v.vector.z = m.vector
            pub.publish(v)


def main(args=None):
    rclpy.init(args=args)
    node = FarnebackCompressedNode()
    try:
        rclpy.spin(node)
    except KeyboardInterrupt:
        pass
    finally:
        node.destroy_node()
        rclpy.shutdown()


if __name__ == '__main__':
    main()
