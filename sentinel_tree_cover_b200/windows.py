"""Integer window bookkeeping of the subtile loop (bit-exact host logic).

  subtile_windows(Lx, Ly, size)  -> (tiles_folder, tiles_array)
     process_subtiles, /root/reference/src/download_and_predict_job.py:1295-1317 and
     make_overlapping_windows, /root/reference/src/tof/tof_downloading.py:498-524.
     tiles_folder rows = (x0, y0, size, size) output placements; tiles_array rows =
     (x0, y0, nx, ny) input slices grown by 7 px per interior side (the model eats
     size+14 and the outer 7 px of border tiles are reflect-padded later, :1378-1395).
  superres_windows(L, wsize)     -> window starts of superresolve_large_tile (:122-123)
"""
import numpy as np


def _starts(L, size, n_rows=6):
    gap = int(np.ceil((L - size) / (n_rows - 1)))
    return np.hstack([np.arange(0, L - size, gap), np.array(L - size)])


def subtile_windows(Lx, Ly, size, n_rows=6, diff=7):
    sx, sy = _starts(Lx, size, n_rows), _starts(Ly, size, n_rows)
    # x-major cartesian product; the reference reaches the same table through a
    # column-wise np.sort + np.tile(np.unique(.)) (:1311-1315), reproduced here
    gx, gy = np.meshgrid(sx, sy)
    folder = np.stack([gx.ravel(), gy.ravel(), np.full(gx.size, size), np.full(gx.size, size)], 1)
    folder = np.sort(folder, axis=0)
    uy = np.unique(folder[:, 1])
    folder[:, 1] = np.tile(uy, int(len(folder) / len(uy)))

    arr = folder.astype(np.int64).copy()
    n_x = int(np.sum(arr[:, 0] == 0))      # windows in the first x group
    n_y = int(np.sum(arr[:, 1] == 0))      # period used for the y-edge test
    i = np.arange(len(arr))
    grow_x = np.full(len(arr), 2 * diff)
    grow_x[:n_x] = diff
    grow_x[-n_x:] = diff
    if 2 * n_x > len(arr):                 # overlapping head/tail groups accumulate like the in-place +=
        grow_x = np.zeros(len(arr), np.int64)
        grow_x[:n_x] += diff
        grow_x[-n_x:] += diff
        grow_x[n_x:-n_x] += 2 * diff
    edge_y = (i % n_y == 0) | ((i + 1) % n_y == 0)
    arr[:, 2] += grow_x
    arr[:, 3] += np.where(edge_y, diff, 2 * diff)
    arr[n_x:, 0] -= diff
    arr[:, 1] -= diff
    arr[arr < 0] = 0
    return folder, arr


def superres_windows(L, wsize=110):
    return [x for x in range(0, L - wsize, wsize)] + [L - wsize]
