"""Camera matrices for the DDA depth-image checker: the inverse view / projection matrices the reference viewer
hands to its shader (src/svviewer/octree_dda_renderer.cpp:393-394, 419-420), float32, column-major flattened."""
from __future__ import annotations

import numpy as np


def look_at_inv(eye, target, up=(0.0, 0.0, 1.0)) -> np.ndarray:
    """Inverse of a right-handed look-at view matrix (camera looks down -Z): camera -> world."""
    eye = np.asarray(eye, np.float64)
    f = np.asarray(target, np.float64) - eye
    f /= np.linalg.norm(f)
    s = np.cross(f, np.asarray(up, np.float64))
    s /= np.linalg.norm(s)
    u = np.cross(s, f)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = s, u, -f, eye
    return np.ascontiguousarray(m.T.astype(np.float32).reshape(16))   # column-major


def perspective_inv(fovy_deg: float, aspect: float, near: float = 0.01, far: float = 100.0) -> np.ndarray:
    f = 1.0 / np.tan(np.radians(fovy_deg) * 0.5)
    p = np.zeros((4, 4))
    p[0, 0], p[1, 1] = f / aspect, f
    p[2, 2], p[2, 3] = (far + near) / (near - far), 2.0 * far * near / (near - far)
    p[3, 2] = -1.0
    return np.ascontiguousarray(np.linalg.inv(p).T.astype(np.float32).reshape(16))


def projection_factor(fovy_deg: float, height: int, pixel_tolerance: float = 1.0) -> float:
    """octree_dda_renderer.cpp:503-507: inv_2tan_half_fovy / (pixelTolerance / screenHeight)."""
    return float((1.0 / (2.0 * np.tan(0.5 * np.radians(fovy_deg)))) / (pixel_tolerance / height))
