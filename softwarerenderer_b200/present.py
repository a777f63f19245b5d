"""Headless presenter: what SDL_UpdateWindowSurface does in the reference's examples
(Box.cpp:204), as a file -- a 0x00RRGGBB colour buffer written as binary PPM (or PNG when Pillow is there)."""
from __future__ import annotations

import numpy as np


def to_rgb8(color: np.ndarray, width: int, height: int) -> np.ndarray:
    c = np.asarray(color, dtype=np.uint32).reshape(height, width)
    return np.stack([(c >> 16) & 255, (c >> 8) & 255, c & 255], axis=-1).astype(np.uint8)


def write_ppm(path: str, color: np.ndarray, width: int, height: int) -> None:
    rgb = to_rgb8(color, width, height)
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (width, height))
        f.write(rgb.tobytes())


def write_image(path: str, color: np.ndarray, width: int, height: int) -> None:
    """PNG / anything Pillow knows when available, else PPM next to the requested name."""
    try:
        from PIL import Image
        Image.fromarray(to_rgb8(color, width, height), "RGB").save(path)
    except ImportError:
        write_ppm(path.rsplit(".", 1)[0] + ".ppm", color, width, height)
