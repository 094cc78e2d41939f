"""Replay ingestion of recorded flights and the velocity Kalman filter (host side, as in the reference).

* `load_ros_yaml`   reads the yaml dumps of ROS messages the reference records (`flight_experiments/hgtData.yaml`,
                    `first_data/vlsData.yaml`, imuData.yaml, camData.yaml: `!!python/object/new:sensor_msgs.msg._Range.Range`
                    with positional `state` lists) WITHOUT the ROS Python classes the reference needs
                    (evaluate_exp.py:8-14, 51-58; of_library.py:327-351): message objects are rebuilt from the slot
                    order of the ROS message definitions.
* `stamps`, `nearest`  the timestamp bookkeeping of evaluate_exp.py:68-80 (seconds relative to the first camera
                    message, nearest IMU / sonar sample per frame), vectorised.
* `imu_samples`     quaternion -> R -> plane normal and gyro rate per frame (evaluate_exp.py:88-95) as the
                    `ofb_imu_sample` records the GPU path consumes.
* `replay_flight`   evaluate_exp.py:77-120 end to end on the device-resident tracker (ofb200.StreamTracker).
* `VelocityKalman`  the 3-state constant-velocity filter of of_module.py:63-76 (cv2.KalmanFilter(3,3,0) with
                    A = B = C = I, float64 as the reference assigns), `predict(control)` / `correct(measurement)` of_module.py:122,152.

Only NumPy/yaml here; the image work happens in libofb200.so through StreamTracker."""
import numpy as np

from . import _lib
from .velocity import quaternion_to_rotation

# slot order of the ROS message definitions the recordings use
ROS_SLOTS = {
    "Header": ("seq", "stamp", "frame_id"),
    "Time": ("secs", "nsecs"),
    "Duration": ("secs", "nsecs"),
    "Range": ("header", "radiation_type", "field_of_view", "min_range", "max_range", "range"),
    "Imu": ("header", "orientation", "orientation_covariance", "angular_velocity", "angular_velocity_covariance",
            "linear_acceleration", "linear_acceleration_covariance"),
    "Quaternion": ("x", "y", "z", "w"),
    "Vector3": ("x", "y", "z"),
    "Point": ("x", "y", "z"),
    "Twist": ("linear", "angular"),
    "TwistStamped": ("header", "twist"),
    "Pose": ("position", "orientation"),
    "PoseStamped": ("header", "pose"),
    "CompressedImage": ("header", "format", "data"),
    "NavSatFix": ("header", "status", "latitude", "longitude", "altitude", "position_covariance",
                  "position_covariance_type"),
    "NavSatStatus": ("status", "service"),
}


class RosMessage(object):
    """A ROS message rebuilt from its yaml dump: attributes by slot name, `_type` = class name."""

    def __init__(self, type_name, values):
        self._type = type_name
        slots = ROS_SLOTS.get(type_name)
        if slots is None:
            slots = tuple("slot%d" % i for i in range(len(values)))
        if len(values) > len(slots):
            raise ValueError("%s: %d values for %d slots" % (type_name, len(values), len(slots)))
        self._slots = slots[:len(values)]
        for k, v in zip(slots, values):
            setattr(self, k, v)

    def __repr__(self):
        return "%s(%s)" % (self._type, ", ".join("%s=%r" % (k, getattr(self, k)) for k in self._slots))


def _make_loader():
    import yaml

    # libyaml's C parser when PyYAML was built with it: the recordings are multi-megabyte files
    class Loader(getattr(yaml, "CSafeLoader", yaml.SafeLoader)):
        pass

    def construct_new(loader, suffix, node):
        # suffix: "sensor_msgs.msg._Range.Range"; node: mapping {state: [...]} (or args/listitems forms)
        name = suffix.rsplit(".", 1)[-1]
        if isinstance(node, yaml.MappingNode):
            m = loader.construct_mapping(node, deep=True)
            values = m.get("state", m.get("args", []))
            if isinstance(values, dict):           # object dumped with a __dict__ state
                msg = RosMessage(name, [])
                msg._slots = tuple(values)
                for k, v in values.items():
                    setattr(msg, k, v)
                return msg
        else:
            values = loader.construct_sequence(node, deep=True)
        return RosMessage(name, list(values))

    def construct_bytes(loader, node):
        import base64
        return base64.b64decode(loader.construct_scalar(node))

    Loader.add_multi_constructor("tag:yaml.org,2002:python/object/new:", construct_new)
    Loader.add_multi_constructor("tag:yaml.org,2002:python/object:", construct_new)
    Loader.add_constructor("tag:yaml.org,2002:binary", construct_bytes)
    Loader.add_constructor("tag:yaml.org,2002:python/tuple", lambda l, n: tuple(l.construct_sequence(n, deep=True)))
    Loader.add_constructor("tag:yaml.org,2002:python/str", lambda l, n: l.construct_scalar(n))
    Loader.add_constructor("tag:yaml.org,2002:python/unicode", lambda l, n: l.construct_scalar(n))
    return Loader


def load_ros_yaml(path_or_stream, limit=None):
    """-> list of RosMessage. Replaces `yaml.load(file('hgtData.yaml'))` (evaluate_exp.py:51-58), which needs the ROS
    message classes on the import path."""
    import yaml
    Loader = _make_loader()
    if hasattr(path_or_stream, "read"):
        data = yaml.load(path_or_stream, Loader=Loader)
    else:
        with open(path_or_stream, "r") as f:
            data = yaml.load(f, Loader=Loader)
    data = list(data or [])
    return data if limit is None else data[:limit]


def stamps(messages, t0_secs=None):
    """evaluate_exp.py:68-75: float(secs - t0.secs) + nsecs / 1e9 per message; t0_secs defaults to the first
    message's own seconds (the script uses the first CAMERA message: pass `cam[0].header.stamp.secs`)."""
    if len(messages) == 0:
        return np.zeros(0)
    if t0_secs is None:
        t0_secs = messages[0].header.stamp.secs
    return np.array([float(m.header.stamp.secs - t0_secs) + float(m.header.stamp.nsecs) / 10 ** 9 for m in messages])


def nearest(values, times):
    """evaluate_exp.py:79-80: np.argmin(np.abs(values - t)) for every t (first minimum on ties, as argmin)."""
    values = np.asarray(values, dtype=np.float64).reshape(-1)
    times = np.asarray(times, dtype=np.float64).reshape(-1)
    if len(values) == 0:
        raise ValueError("no samples to associate")
    return np.abs(values[None, :] - times[:, None]).argmin(axis=1)


def imu_samples(imu_msgs, range_msgs, imu_idx, hgt_idx, translation=(0.0, 0.0, 1.0)):
    """evaluate_exp.py:82-95 per frame: d = sonar range, n = R(q) e_z, omega = gyro rate, t = `Translation`
    (evaluate_exp.py:48). -> _lib.IMU_DTYPE array."""
    out = np.zeros(len(imu_idx), _lib.IMU_DTYPE)
    for k, (ii, hi) in enumerate(zip(imu_idx, hgt_idx)):
        m = imu_msgs[int(ii)]
        q = m.orientation
        R = quaternion_to_rotation(q.x, q.y, q.z, q.w)
        out["d"][k] = range_msgs[int(hi)].range
        out["n"][k] = R[:, 2]
        out["w"][k] = [m.angular_velocity.x, m.angular_velocity.y, m.angular_velocity.z]
        out["t"][k] = translation
    return out


def replay_flight(frames, cam_times, imu_msgs, imu_times, range_msgs, range_times, tracker, translation=(0.0, 0.0, 1.0)):
    """evaluate_exp.py:77-120 on the GPU tracker: for every frame (grey or BGR array, see tracker.bgr) the nearest
    IMU and sonar samples, then track / filter / solve / top-up. Yields (frame index, result record, imu sample)."""
    cam_times = np.asarray(cam_times, dtype=np.float64)
    ii, hi = nearest(imu_times, cam_times), nearest(range_times, cam_times)
    samples = imu_samples(imu_msgs, range_msgs, ii, hi, translation)
    for k, frame in enumerate(frames):
        res = tracker.step(frame, samples[k:k + 1])
        yield k, res[0], samples[k]


class VelocityKalman(object):
    """of_module.py:63-76: cv2.KalmanFilter(3, 3, 0) with transition = control = measurement = np.eye(3),
    processNoiseCov = 1e-5 I, measurementNoiseCov = 1e1 I, errorCovPost = 0.1 I, statePost = 0. The reference assigns
    float64 matrices, which cv2 then computes in (probed on cv2 4.13), and although the filter is created with
    controlParams = 0 the assigned controlMatrix IS applied by predict(control) (of_module.py:122). Both kept."""

    def __init__(self, process_noise=1e-5, measurement_noise=1e1, error_cov=0.1, dtype=np.float64):
        I = np.eye(3, dtype=dtype)
        self.dtype = dtype
        self.A, self.B, self.C = I.copy(), I.copy(), I.copy()
        self.Q = process_noise * I
        self.R = measurement_noise * I
        self.P_post = error_cov * I
        self.x_post = np.zeros(3, dtype)
        self.x_pre, self.P_pre = self.x_post.copy(), self.P_post.copy()

    def predict(self, control=None):
        x = self.A @ self.x_post
        if control is not None:
            x = x + self.B @ np.asarray(control, self.dtype).reshape(3)
        self.x_pre = x
        self.P_pre = self.A @ self.P_post @ self.A.T + self.Q
        # cv2 copies the prediction into the posterior so that predict() can be called repeatedly
        self.x_post, self.P_post = self.x_pre.copy(), self.P_pre.copy()
        return self.x_pre.reshape(3, 1).copy()

    def correct(self, measurement):
        z = np.asarray(measurement, self.dtype).reshape(3)
        S = self.C @ self.P_pre @ self.C.T + self.R
        K = np.linalg.solve(S, self.C @ self.P_pre).T            # gain = P- C^T S^-1 (S symmetric)
        self.x_post = self.x_pre + K @ (z - self.C @ self.x_pre)
        self.P_post = self.P_pre - K @ self.C @ self.P_pre
        return self.x_post.reshape(3, 1).copy()
