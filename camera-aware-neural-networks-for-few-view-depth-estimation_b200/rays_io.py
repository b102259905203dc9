"""Ray-direction .bin codec: int32 H, int32 W, then H*W*3 float32 row-major
(reference: src/preprocessing/ray_direction_computer.h:96-99, ray_direction_computer.cpp:129-201;
the loader reshapes it to (3,H,W): src/data/sunrgbd_loader.cpp:345-347)."""
from __future__ import annotations

import numpy as np


def save_ray_directions(rays_hw3, height: int, width: int, filename: str) -> bool:
    a = np.ascontiguousarray(np.asarray(rays_hw3, dtype=np.float32))
    if a.size != height * width * 3:
        return False     # dimension mismatch -> false, like ray_direction_computer.cpp:146-152
    with open(filename, "wb") as f:
        np.array([height, width], dtype="<i4").tofile(f)
        a.reshape(-1).astype("<f4").tofile(f)
    return True


def load_ray_directions(filename: str):
    """Returns (rays (H*W,3) float32, H, W)."""
    with open(filename, "rb") as f:
        hdr = np.fromfile(f, dtype="<i4", count=2)
        if hdr.size != 2:
            raise RuntimeError(f"Error: Could not open file for reading: {filename}")
        h, w = int(hdr[0]), int(hdr[1])
        data = np.fromfile(f, dtype="<f4", count=h * w * 3)
    if data.size != h * w * 3:
        raise RuntimeError(f"truncated ray file: {filename}")
    return data.reshape(h * w, 3), h, w
