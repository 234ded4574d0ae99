import numpy as np
def view_matrix(eye, yaw, pitch):
    """row-major view matrix looking along +z of the rotated camera frame (the reference's projection has w = +z)"""
    cy, sy, cp, sp = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch)
    Ry = np.array([[cy, 0, -sy, 0], [0, 1, 0, 0], [sy, 0, cy, 0], [0, 0, 0, 1]], np.float64)
    Rx = np.array([[1, 0, 0, 0], [0, cp, sp, 0], [0, -sp, cp, 0], [0, 0, 0, 1]], np.float64)
    T = np.eye(4); T[:3, 3] = -np.asarray(eye, np.float64)
    return (Rx @ Ry @ T).astype(np.float32)
def projection(fov_deg, aspect, n, f):
    """Matrix::projection (math/src/Matrix.cpp:378-399), restated for test inputs only"""
    scale = np.float32(np.tan(fov_deg * 0.5 * np.float32(np.float32(3.14159265) / np.float32(180.0)))) * np.float32(n)
    r = np.float32(aspect) * scale; l = -r; t = scale; b = -t
    P = np.zeros((4, 4), np.float32)
    P[0, 0] = 2 * n / (r - l); P[0, 2] = -(r + l) / (r - l)
    P[1, 1] = 2 * n / (t - b); P[1, 2] = -(t + b) / (t - b)
    P[2, 2] = f / (f - n); P[2, 3] = -f * n / (f - n)
    P[3, 2] = 1
    return P
