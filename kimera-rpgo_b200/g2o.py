"""g2o ingest for the harness.

Mirrors what gtsam::load3D / load2D hand to RobustSolver::update in the reference's examples
(examples/RpgoReadG2o.cpp:129-143): VERTEX_SE3:QUAT / EDGE_SE3:QUAT (and VERTEX_SE2 / EDGE_SE2)
become values and BetweenFactors whose covariance is the inverse of the g2o information matrix
re-ordered from g2o's [translation; rotation] to GTSAM's [rotation; translation] tangent order.
Pure host-side parsing; nothing here touches the GPU.
"""
import numpy as np


def quat_to_R(qx, qy, qz, qw):
    n = np.sqrt(qx * qx + qy * qy + qz * qz + qw * qw)
    x, y, z, w = qx / n, qy / n, qz / n, qw / n
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _upper_tri(vals, n):
    m = np.zeros((n, n))
    k = 0
    for i in range(n):
        for j in range(i, n):
            m[i, j] = m[j, i] = vals[k]
            k += 1
    return m


def load3d(path):
    """Returns (values, edges): values = [(key, pose12)], edges = [(key1, key2, pose12, cov6x6)] in file order."""
    values, edges = [], []
    with open(path) as f:
        for line in f:
            p = line.split()
            if not p:
                continue
            if p[0] == "VERTEX_SE3:QUAT":
                key = int(p[1])
                v = [float(x) for x in p[2:9]]
                R = quat_to_R(v[3], v[4], v[5], v[6])
                values.append((key, np.concatenate([R.reshape(9), v[:3]])))
            elif p[0] == "EDGE_SE3:QUAT":
                k1, k2 = int(p[1]), int(p[2])
                v = [float(x) for x in p[3:10]]
                R = quat_to_R(v[3], v[4], v[5], v[6])
                m = _upper_tri([float(x) for x in p[10:31]], 6)
                g = np.zeros((6, 6))
                g[0:3, 0:3] = m[3:6, 3:6]
                g[3:6, 3:6] = m[0:3, 0:3]
                g[0:3, 3:6] = m[0:3, 3:6]
                g[3:6, 0:3] = m[3:6, 0:3]
                cov = np.linalg.inv(g)
                edges.append((k1, k2, np.concatenate([R.reshape(9), v[:3]]), cov))
    return values, edges


def load2d(path):
    """Returns (values, edges) with pose = (cos, sin, x, y) and 3x3 covariance in (x, y, theta) order."""
    values, edges = [], []
    with open(path) as f:
        for line in f:
            p = line.split()
            if not p:
                continue
            if p[0] == "VERTEX_SE2":
                key = int(p[1])
                x, y, th = (float(v) for v in p[2:5])
                values.append((key, np.array([np.cos(th), np.sin(th), x, y])))
            elif p[0] == "EDGE_SE2":
                k1, k2 = int(p[1]), int(p[2])
                x, y, th = (float(v) for v in p[3:6])
                m = _upper_tri([float(v) for v in p[6:12]], 3)
                edges.append((k1, k2, np.array([np.cos(th), np.sin(th), x, y]), np.linalg.inv(m)))
    return values, edges


def R_to_quat(R):
    """rotation matrix -> (qx, qy, qz, qw), w >= 0 branch-stable (Shepperd)"""
    R = np.asarray(R, dtype=np.float64).reshape(3, 3)
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    if tr > 0:
        s = np.sqrt(tr + 1.0) * 2
        w, x, y, z = 0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = np.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        w, x, y, z = (R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s
    elif R[1, 1] > R[2, 2]:
        s = np.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        w, x, y, z = (R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s
    else:
        s = np.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        w, x, y, z = (R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s
    return x, y, z, w


def write_g2o(path, values, edges, d=3):
    """Egress in the record layout of the reference's writeG2o (src/Logger.cpp:64-163): VERTEX_SE3:QUAT / VERTEX_SE2
    lines for the values, then EDGE_SE3:QUAT / EDGE_SE2 lines with the upper triangle of the information matrix in
    g2o order (translation first for SE3).  edges: (key1, key2, pose, cov); numbers are written with 17 significant
    digits so that a round trip through load3d / load2d is lossless (the reference uses the stream default)."""
    f17 = lambda v: "%.17g" % v
    with open(path, "w") as f:
        for key, p in values:
            if d == 3:
                q = R_to_quat(np.asarray(p[:9]))
                f.write("VERTEX_SE3:QUAT %d %s\n" % (key, " ".join(f17(v) for v in list(p[9:12]) + list(q))))
            else:
                f.write("VERTEX_SE2 %d %s\n" % (key, " ".join(f17(v) for v in (p[2], p[3], np.arctan2(p[1], p[0])))))
        for k1, k2, p, cov in edges:
            info = np.linalg.inv(np.asarray(cov, dtype=np.float64))
            if d == 3:
                q = R_to_quat(np.asarray(p[:9]))
                g = np.zeros((6, 6))
                g[0:3, 0:3] = info[3:6, 3:6]
                g[3:6, 3:6] = info[0:3, 0:3]
                g[0:3, 3:6] = info[0:3, 3:6]
                g[3:6, 0:3] = info[3:6, 0:3]
                tri = [g[i, j] for i in range(6) for j in range(i, 6)]
                f.write("EDGE_SE3:QUAT %d %d %s\n" % (k1, k2, " ".join(f17(v) for v in list(p[9:12]) + list(q) + tri)))
            else:
                tri = [info[i, j] for i in range(3) for j in range(i, 3)]
                f.write("EDGE_SE2 %d %d %s\n" % (k1, k2, " ".join(f17(v) for v in [p[2], p[3], np.arctan2(p[1], p[0])] + tri)))
