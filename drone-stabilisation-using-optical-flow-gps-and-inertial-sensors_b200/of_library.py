"""of_library-compatible module (reference: of_library.py, 15 functions). Import it in place of the
reference's `of_library`:   import ofb200.of_library as of

Hot-path members are GPU-backed through libofb200.so:
    pix_trans   of_library.py:31-43      (host arithmetic on two integers; also fused into frame_pairs)
    r_tilde     of_library.py:365-386    (5-arg) and the 4-arg homogeneous copy
                sensor_precision_experiments/pixhawk_pure_IMU/of_library.py:365-384
    initialize_ft of_library.py:231-263  (the intent: gftt -> LK loop -> immobile filter -> ranking)
The remaining helpers are element-wise NumPy / cv2 drawing utilities that the reference itself runs on
the host; several of them do not execute in the reference (NameErrors: convert_to_of, dynamic_immobile,
calc_height, eval_ft, initialize_ft, distancecluster) -- here they implement the evident intent.
"""
import numpy as np

from . import _lib
from . import vision


def visualize(image, mask, newpos, oldpos, frame_name="visualization", marker=[0, 0, 255]):
    """of_library.py:19-27 (GUI; needs an OpenCV build with highgui)."""
    import cv2
    img = image
    for new, old in zip(newpos, oldpos):
        a, b = (int(v) for v in np.ravel(new)[:2])
        c, d = (int(v) for v in np.ravel(old)[:2])
        mask = cv2.line(mask, (a, b), (c, d), marker, 2)
        image = cv2.circle(image, (a, b), 5, marker, -1)
        img = cv2.add(image, mask)
        cv2.imshow(frame_name, img)
    return img


def pix_trans(img_dim):
    """of_library.py:31-43: principal point, d/2 for even d else (d+1)/2 per axis."""
    trans_x = img_dim[0] / 2 if img_dim[0] % 2 == 0 else (img_dim[0] + 1) / 2
    trans_y = img_dim[1] / 2 if img_dim[1] % 2 == 0 else (img_dim[1] + 1) / 2
    return trans_x, trans_y


def convert_to_of(pos, pos_err, speed, speed_err, height, height_err, focal_len, img_dim):
    """of_library.py:53-75 (expected flow of a static point under the pin-hole model + variance)."""
    eps = np.finfo(float).eps
    height = np.asarray(height, dtype=float)
    if np.any(height < eps):
        raise ValueError(' height over feature is Zero or Negative')
    trans_x, trans_y = pix_trans(img_dim)
    pos = np.asarray(pos, dtype=float)
    pos_err = np.asarray(pos_err, dtype=float)
    x_exp = (focal_len - (pos[0, :] - trans_x) / height) * speed[0] / height
    y_exp = (focal_len - (pos[1, :] - trans_y) / height) * speed[1] / height
    x_err = ((pos_err[0, :] * speed[0] / height) ** 2 + ((focal_len - pos[0, :] + trans_x) * speed_err[0] / height) ** 2
             + ((focal_len - pos[0, :] + trans_x) * speed[0] * height_err / height ** 2) ** 2)
    y_err = ((pos_err[1, :] * speed[1] / height) ** 2 + ((focal_len - pos[1, :] + trans_y) * speed_err[1] / height) ** 2
             + ((focal_len - pos[1, :] + trans_y) * speed[1] * height_err / height ** 2) ** 2)
    return [x_exp, y_exp], [x_err, y_err]


def static_immobile(newpos, oldpos, maxspeed, distance, dummy_value):
    """of_library.py:88-92."""
    speed_constraint = (np.abs(newpos - oldpos)) < (maxspeed / distance)
    dummy_constraint = (oldpos) != dummy_value
    stable = speed_constraint * dummy_constraint
    return stable[:, :, 0] * stable[:, :, 1]


def dynamic_immobile(newpos, newpos_err, oldpos, oldpos_err, speed, speed_err, focal_len, dummy_value, height,
                     height_err, img_dim):
    """of_library.py:100-114."""
    newpos = np.asarray(newpos, dtype=float)
    oldpos = np.asarray(oldpos, dtype=float)
    of_obs = newpos - oldpos
    of_obs_err = np.asarray(oldpos_err, dtype=float) ** 2 + np.asarray(newpos_err, dtype=float) ** 2
    p = newpos.reshape(-1, 2).T
    pe = np.broadcast_to(np.asarray(newpos_err, dtype=float).reshape(len(p.T), -1), (len(p.T), 2)).T
    of_exp, of_exp_err = convert_to_of(p, pe, speed, speed_err, height, height_err, focal_len, img_dim)
    of_exp = np.stack(of_exp, axis=-1).reshape(newpos.shape)
    of_exp_err = np.stack(of_exp_err, axis=-1).reshape(newpos.shape)
    speed_constraint = ((of_obs - of_exp) ** 2) < (np.broadcast_to(of_obs_err.reshape(len(newpos), 1, -1), newpos.shape)
                                                   + of_exp_err)
    dummy_constraint = (oldpos) != dummy_value
    stable = speed_constraint * dummy_constraint
    return stable[:, :, 0] * stable[:, :, 1]


def kmeancluster(points, k):
    """of_library.py:121-136 (list of per-cluster arrays; the reference's ragged np.array fails on numpy>=1.24)."""
    import cv2
    criteria = (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_MAX_ITER, 10, 1.0)
    points = np.float32(points)
    ret, label, center = cv2.kmeans(points, k, None, criteria, 10, cv2.KMEANS_RANDOM_CENTERS)
    out = np.empty(k, dtype=object)
    for i in range(k):
        out[i] = points[label.ravel() == i]
    return out


def distancecluster(pointcloud, points, maxdist, clusterlist):
    """of_library.py:146-171: single-linkage growth of index clusters (L-inf distance < maxdist)."""
    pointcloud = np.asarray(pointcloud, dtype=float).reshape(-1, 2)
    points = np.asarray(points, dtype=float).reshape(-1, 2)
    clusterlist = [list(np.atleast_1d(c)) for c in clusterlist]
    for i in range(len(points)):
        cloud = np.vstack([pointcloud, points[:i]]) if i else pointcloud
        near = np.where(np.all(np.abs(cloud - points[i]) < maxdist, axis=1))[0] if len(cloud) else np.array([], int)
        new_index = len(pointcloud) + i
        fuse = [ci for ci, c in enumerate(clusterlist) if np.isin(c, near).any()]
        merged = [new_index]
        for ci in fuse:
            merged = list(clusterlist[ci]) + merged
        clusterlist = [c for ci, c in enumerate(clusterlist) if ci not in fuse]
        clusterlist.append(merged)
    return clusterlist, np.vstack([pointcloud, points]) if len(points) else pointcloud


def circles(points, mask, radius):
    """of_library.py:204-208 (the reference uses point[0] for y as well; the fixed copy in
    sensor_precision_experiments uses point[1] -- that one is followed)."""
    import cv2
    for point in points:
        cv2.circle(mask, (int(point[0]), int(point[1])), radius, 0, cv2.FILLED)


def boundingboxes(clusterlist, mask, radius):
    """of_library.py:184-193."""
    import cv2
    for cluster in clusterlist:
        cluster = np.asarray(cluster, dtype=np.float32)
        if len(cluster) > 1:
            box = np.intp(cv2.boxPoints(cv2.minAreaRect(cluster.reshape(-1, 1, 2))))
            cv2.drawContours(mask, [box], 0, 0, cv2.FILLED)
        elif len(cluster) == 1:
            circles(cluster.reshape(-1, 2), mask, radius)


def convexhull(clusterlist, mask, radius):
    """of_library.py:219-226."""
    import cv2
    for cluster in clusterlist:
        cluster = np.asarray(cluster, dtype=np.float32)
        if len(cluster) > 1:
            filler = np.array(cv2.convexHull(cluster.reshape(-1, 1, 2), returnPoints=True), dtype='int32')
            cv2.fillConvexPoly(mask, filler, 0)
        elif len(cluster) == 1:
            circles(cluster.reshape(-1, 2), mask, radius)


def calc_height(of, of_err, vel, vel_err, focal_len, newpos, newpos_err):
    """of_library.py:270-286: per-feature height from the pin-hole model, both axes averaged."""
    of = np.asarray(of, dtype=float).reshape(-1, 2)
    of_err = np.broadcast_to(np.asarray(of_err, dtype=float).reshape(len(of), -1), of.shape)
    newpos = np.asarray(newpos, dtype=float).reshape(-1, 2)
    newpos_err = np.broadcast_to(np.asarray(newpos_err, dtype=float).reshape(len(of), -1), of.shape)
    vel = np.broadcast_to(np.asarray(vel, dtype=float).reshape(-1, 3), (len(of), 3))
    vel_err = np.broadcast_to(np.asarray(vel_err, dtype=float).reshape(-1, 3), (len(of), 3))
    h, he = [], []
    for a in (0, 1):
        num = focal_len * vel[:, a] - newpos[:, a] * vel[:, 2]
        h.append(num / of[:, a])
        he.append((focal_len * vel_err[:, a] / of[:, a]) ** 2 + (num * of_err[:, a] / of[:, a] ** 2) ** 2
                  + (newpos_err[:, a] * vel[:, 2] / of[:, a]) ** 2 + (newpos[:, a] * vel_err[:, 2] / of[:, a]) ** 2)
    return 0.5 * (h[0] + h[1]), he[0] + he[1]


def eval_ft(weight, height, height_err, new_pos, new_pos_err, img_dim):
    """of_library.py:291-317: rank features by a weighted score; returns the four arrays sorted."""
    def norm(a):
        a = np.asarray(a, dtype=float)
        span = np.amax(a) - np.amin(a)
        return (a - np.amin(a)) / span if span > 0 else np.zeros_like(a)
    height = np.asarray(height, dtype=float)
    height_err = np.asarray(height_err, dtype=float)
    pos = np.asarray(new_pos, dtype=float).reshape(-1, 2)
    perr = np.asarray(new_pos_err, dtype=float).reshape(len(pos), -1).mean(axis=1)
    trans = pix_trans(img_dim)
    quad = (pos[:, 0] - trans[0]) ** 2 + (pos[:, 1] - trans[1]) ** 2
    dist_norm = quad / np.amax(quad) if np.amax(quad) > 0 else np.zeros_like(quad)
    best = weight[0] * (1 - norm(height)) + weight[1] * norm(height_err) + weight[2] * (1 - dist_norm) + weight[3] * norm(perr)
    idx = best.argsort()
    return height[idx], height_err[idx], np.asarray(new_pos)[idx], np.asarray(new_pos_err)[idx]


def initialize_ft(camera, feature_parameter, lk_parameter, iterations, end_count, vel, vel_err, focal_len, dummy_value,
                  img_dim, weight):
    """of_library.py:231-263, as intended: first frame -> goodFeaturesToTrack -> `iterations` LK steps with
    the dynamic immobile filter -> eval_ft ranking. `camera` is anything cv2.VideoCapture accepts, or an
    iterable of BGR / grey frames."""
    if end_count <= 0:
        raise ValueError(' end_count must be a positive number')
    if iterations <= 0:
        raise ValueError(' iterations must be a positive number')
    frames = _frame_source(camera)
    old_gray = _to_gray(next(frames))
    old_pos = vision.goodFeaturesToTrack(old_gray, mask=None, **feature_parameter)
    if old_pos is None:
        raise ValueError(' no features found in the first frame')
    old_pos_err = np.zeros((len(old_pos), 1), np.float32)
    height = height_err = new_pos = new_pos_err = None
    for i in range(iterations):
        frame_gray = _to_gray(next(frames))
        new_pos, status, new_pos_err = vision.calcOpticalFlowPyrLK(old_gray, frame_gray, old_pos, None, **lk_parameter)
        if i == 0:
            height, height_err = calc_height(new_pos - old_pos, new_pos_err, vel, vel_err, focal_len, new_pos, new_pos_err)
        keep = dynamic_immobile(new_pos, new_pos_err.reshape(-1, 1, 1), old_pos, old_pos_err.reshape(-1, 1, 1), vel, vel_err,
                                focal_len, dummy_value, height, height_err, img_dim).reshape(-1) * status.reshape(-1)
        keep = keep.astype(bool)
        old_pos = new_pos[keep].reshape(-1, 1, 2)
        old_pos_err = new_pos_err[keep]
        height, height_err = height[keep], height_err[keep]
        new_pos, new_pos_err = old_pos, old_pos_err
        old_gray = frame_gray
        if len(old_pos) == 0:
            break
    if len(old_pos) == 0:
        return np.array([]), np.array([]), old_pos, old_pos_err
    return eval_ft(weight, height, height_err, new_pos, new_pos_err, img_dim)


def _frame_source(camera):
    if isinstance(camera, (str, int)):
        import cv2
        cap = cv2.VideoCapture(camera)

        def gen():
            while True:
                ret, frame = cap.read()
                if not ret:
                    return
                yield frame
        return gen()
    return iter(camera)


def _to_gray(frame):
    frame = np.asarray(frame)
    return vision.cvtColor(frame) if frame.ndim == 3 else np.ascontiguousarray(frame, dtype=np.uint8)


def read_yaml_imu(yamlfile):
    """of_library.py:327-351. The dump is parsed with replay.load_ros_yaml: a SafeLoader that rebuilds the ROS
    messages as inert attribute bags, so neither the ROS classes nor PyYAML's object-constructing loader are needed."""
    from .replay import load_ros_yaml
    imuData = load_ros_yaml(yamlfile)
    datastack = []
    for entry in reversed(imuData):
        o, la, av = entry.orientation, entry.linear_acceleration, entry.angular_velocity
        datastack.append([entry.header.stamp.secs + float(entry.header.stamp.nsecs / 10 ** 6),
                          [o.x, o.y, o.z, o.w], entry.orientation_covariance,
                          [la.x, la.y, la.z], entry.linear_acceleration_covariance,
                          [av.x, av.y, av.z], entry.angular_velocity_covariance])
    return datastack


def r_tilde(x, u, n, v, dist=None, ctx=None):
    """of_library.py:365-386: cos of the angle between X x v and X x u, and the implied distance.
    r_tilde(x,u,n,v,dist) with (N,2) x,u; r_tilde(x,u,n,v) with homogeneous (N,3) x,u (older copy)."""
    ctx = ctx or _lib.default_context()
    x = np.asarray(x, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    if x.ndim == 3 and x.shape[1] == 1:
        x = x.reshape(len(x), -1)
        u = u.reshape(len(u), -1)
    if dist is None:
        if x.ndim != 2 or x.shape[1] != 3 or u.shape != x.shape:
            raise ValueError("the 4-argument r_tilde expects homogeneous (N,3) x and u")
        ld, dd = 3, 0.0
    else:
        if x.ndim != 2 or x.shape[1] != 2 or u.shape != x.shape:
            raise ValueError("r_tilde(x,u,n,v,dist) expects (N,2) x and u")
        ld, dd = 2, float(dist)
    x = np.ascontiguousarray(x)
    u = np.ascontiguousarray(u)
    n3 = np.ascontiguousarray(np.asarray(n, dtype=np.float64).reshape(3))
    v3 = np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(3))
    r = np.zeros(len(x))
    d = np.ones(len(x))
    _lib.check(ctx.lib.ofb_r_tilde(ctx.h, _lib.ptr(x), _lib.ptr(u), len(x), ld, _lib.ptr(n3), _lib.ptr(v3), dd,
                                   _lib.ptr(r), _lib.ptr(d)))
    return r, d
