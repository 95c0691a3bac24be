"""RGB frames of the whole grid (MiniGridEnv.get_full_render) for selected envs of a batch.

Reference: gym_minigrid/minigrid.py:1514-1563 (get_full_render: highlight mask = the agent's view
window), :712-747 (Grid.render: one cached tile image per cell), :662-710 (render_tile: grid lines, the
object, the agent triangle, highlight, 3x3 supersampling) and gym_minigrid/rendering.py (the
primitives).  A tile image depends only on (cell type, agent on it, highlighted, tile size) -- the
reference caches exactly that -- so the 16 possible tiles are rasterised once on the host (`tile_atlas`,
a few thousand point tests) and a CUDA kernel (ta_render) blits them into [M, 17*ts, 17*ts, 3] frames:
the per-frame work is a pure HBM-bound copy.
"""
from __future__ import annotations

import ctypes as C
import math
from functools import lru_cache
from typing import Optional

import numpy as np
import torch

from . import _capi

GRID_LINE = (100, 100, 100)          # render_tile, minigrid.py:681-682
CELL_COLOR = {1: (100, 100, 100), 2: (255, 255, 0), 3: (0, 255, 0)}   # wall grey, ball yellow, goal green
SUBDIVS = 3


def _centres(n):
    c = (np.arange(n) + 0.5) / n      # rendering.fill_coords: (x + 0.5) / width
    return np.meshgrid(c, c)          # xf varies along columns, yf along rows


def _rect(xf, yf, xmin, xmax, ymin, ymax):
    return (xf >= xmin) & (xf <= xmax) & (yf >= ymin) & (yf <= ymax)


def _triangle(xf, yf, a, b, c):
    """rendering.point_in_triangle: barycentric test with float32 corner arithmetic."""
    a, b, c = (np.array(v, dtype=np.float32) for v in (a, b, c))
    v0, v1 = c - a, b - a
    v2x, v2y = xf - a[0], yf - a[1]                       # float64 - float32 -> float64
    dot00 = v0[0] * v0[0] + v0[1] * v0[1]
    dot01 = v0[0] * v1[0] + v0[1] * v1[1]
    dot11 = v1[0] * v1[0] + v1[1] * v1[1]
    dot02 = v0[0] * v2x + v0[1] * v2y
    dot12 = v1[0] * v2x + v1[1] * v2y
    inv = 1 / (dot00 * dot11 - dot01 * dot01)
    u = (dot11 * dot02 - dot01 * dot12) * inv
    v = (dot00 * dot12 - dot01 * dot02) * inv
    return (u >= 0) & (v >= 0) & ((u + v) < 1)


def _tile(code: int, agent: bool, highlight: bool, ts: int) -> np.ndarray:
    n = ts * SUBDIVS
    img = np.zeros((n, n, 3), np.uint8)
    xf, yf = _centres(n)
    img[_rect(xf, yf, 0, 0.031, 0, 1)] = GRID_LINE
    img[_rect(xf, yf, 0, 1, 0, 0.031)] = GRID_LINE
    if code in (1, 3):                                    # Wall / Goal: the whole tile
        img[_rect(xf, yf, 0, 1, 0, 1)] = CELL_COLOR[code]
    elif code == 2:                                       # Ball: circle of radius 0.31
        img[(xf - 0.5) * (xf - 0.5) + (yf - 0.5) * (yf - 0.5) <= 0.31 * 0.31] = CELL_COLOR[2]
    if agent:                                             # red triangle, rotated by agent_dir = 3
        theta = 0.5 * math.pi * 3
        x, y = xf - 0.5, yf - 0.5
        x2 = 0.5 + x * math.cos(-theta) - y * math.sin(-theta)
        y2 = 0.5 + y * math.cos(-theta) + x * math.sin(-theta)
        img[_triangle(x2, y2, (0.12, 0.19), (0.87, 0.50), (0.12, 0.81))] = (255, 0, 0)
    if highlight:                                         # rendering.highlight_img, alpha 0.3 towards white
        blend = img + 0.30 * (np.array((255, 255, 255), dtype=np.uint8) - img)
        img = blend.clip(0, 255).astype(np.uint8)
    small = img.reshape(ts, SUBDIVS, ts, SUBDIVS, 3).mean(axis=3).mean(axis=1)
    return small.astype(np.uint8)                          # Grid.render stores the float tile into a uint8 image


@lru_cache(maxsize=8)
def tile_atlas(tile_size: int) -> np.ndarray:
    """uint8 [16, ts, ts, 3]; index = cell code | agent << 2 | highlighted << 3."""
    out = np.zeros((16, tile_size, tile_size, 3), np.uint8)
    for idx in range(16):
        out[idx] = _tile(idx & 3, bool(idx & 4), bool(idx & 8), tile_size)
    return out


def compose(grid_codes: np.ndarray, agent_xy, tile_size: int, highlight: bool, view: int) -> np.ndarray:
    """Host composition of one frame (what the kernel does per env): grid_codes [289] index y*17+x."""
    atlas = tile_atlas(tile_size)
    ax, ay = agent_xy
    img = np.zeros((17 * tile_size, 17 * tile_size, 3), np.uint8)
    for y in range(17):
        for x in range(17):
            hl = highlight and (ax - view // 2 <= x <= ax + view // 2) and (ay - (view - 1) <= y <= ay)
            idx = int(grid_codes[y * 17 + x]) | (4 if (x, y) == (ax, ay) else 0) | (8 if hl else 0)
            img[y * tile_size:(y + 1) * tile_size, x * tile_size:(x + 1) * tile_size] = atlas[idx]
    return img


_atlas_dev = {}


def render(env, env_ids: Optional[torch.Tensor] = None, tile_size: int = 32, highlight: bool = False) -> torch.Tensor:
    """get_full_render() of the envs `env_ids` (int64, default all) of a TwoarmyVecEnv:
    uint8 [M, 17*ts, 17*ts, 3] on the env's device."""
    key = (str(env.device), int(tile_size))
    if key not in _atlas_dev:
        _atlas_dev[key] = torch.from_numpy(tile_atlas(int(tile_size))).to(env.device).contiguous()
    atlas = _atlas_dev[key]
    if env_ids is not None:
        env_ids = env_ids.to(device=env.device, dtype=torch.int64).contiguous()
        m = env_ids.numel()
    else:
        m = env.num_envs
    side = 17 * int(tile_size)
    out = torch.empty((m, side, side, 3), dtype=torch.uint8, device=env.device)
    st = C.c_void_p(torch.cuda.current_stream(env.device).cuda_stream)
    _capi.check(_capi.lib().ta_render(env._h, C.c_void_p(atlas.data_ptr()), int(tile_size), int(bool(highlight)),
                                      None if env_ids is None else C.c_void_p(env_ids.data_ptr()), m,
                                      C.c_void_p(out.data_ptr()), st), "ta_render")
    return out
