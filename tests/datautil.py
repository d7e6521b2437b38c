"""Readers for the reference's text formats (mapParam / mapValue / Lidar), used by tests and tools.

Formats: SURVEY.md §8b — mapParam.txt = `cols rows resol oriX oriY`; mapValue*.txt = rows x cols ints stored
through `%d` into a uint8 slot (LSD/main_on_windows.cpp:43-45), i.e. value & 0xFF."""
import numpy as np


def load_map_param(path):
    v = np.loadtxt(path).reshape(-1)
    return dict(cols=int(v[0]), rows=int(v[1]), res=float(v[2]), ori_x=float(v[3]), ori_y=float(v[4]))


def load_map_value(path, cols, rows):
    a = np.fromfile(path, dtype=np.int64, sep=" ")
    assert a.size == cols * rows, (a.size, cols, rows)
    return (a & 0xFF).astype(np.uint8).reshape(rows, cols)


def load_lidar_frames(path, per_loop=360):
    """Lidar.txt: `range angle` per line, 360 lines per frame; Inf ranges dropped (main_on_windows.cpp:110-123)."""
    a = np.loadtxt(path).reshape(-1, 2)
    nf = len(a) // per_loop
    frames = []
    for f in range(nf):
        fr = a[f * per_loop:(f + 1) * per_loop]
        keep = np.isfinite(fr[:, 0])
        frames.append((fr[keep, 0].copy(), fr[keep, 1].copy()))
    return frames
