"""Synthetic benchmark inputs: particles uniform over free cells and 360-beam scans
ray-cast from the map.  Input generation only -- not part of the hot path."""
import numpy as np


def free_space_particles(gm, n, seed=1234):
    """x, y uniform inside uniformly chosen free cells, theta ~ U[-pi, pi); (n, 3) float64."""
    rs = np.random.RandomState(seed)
    free = np.flatnonzero(gm.occ.ravel() == 0)
    cells = free[rs.randint(0, len(free), n)]
    my, mx = np.divmod(cells, gm.width)
    x = gm.origin_x + (mx + rs.uniform(0, 1, n)) * gm.resolution
    y = gm.origin_y + (my + rs.uniform(0, 1, n)) * gm.resolution
    th = rs.uniform(-np.pi, np.pi, n)
    return np.column_stack((x, y, th))


def raycast_scan(gm, pose, num_beams=360, sensor_max=3.5, step_size=0.1, noise_sigma=0.0, seed=4321):
    """Ray-marching range scan in the style of the reference's raycast (pu:4-29): 0.1 m steps, first
    cell with occupancy != 0 (or leaving the map) ends the ray; >= sensor_max -> +inf."""
    angles = np.linspace(0.0, 2 * np.pi - 2 * np.pi / num_beams, num_beams, dtype=np.float32)
    blocked = gm.occ != 0
    a = pose[2] + angles.astype(np.float64)
    dx, dy = np.cos(a), np.sin(a)
    ranges = np.full(num_beams, sensor_max)
    alive = np.ones(num_beams, bool)
    for i in range(1, int(sensor_max / step_size) + 1):
        cx = pose[0] + i * step_size * dx
        cy = pose[1] + i * step_size * dy
        gx = ((cx - gm.origin_x) / gm.resolution).astype(np.int64)
        gy = ((cy - gm.origin_y) / gm.resolution).astype(np.int64)
        inside = (gx >= 0) & (gx < gm.width) & (gy >= 0) & (gy < gm.height)
        hit = np.zeros(num_beams, bool)
        hit[inside] = blocked[gy[inside], gx[inside]]
        stop_out = alive & ~inside
        stop_hit = alive & inside & hit
        ranges[stop_hit] = i * step_size
        alive &= ~(stop_out | stop_hit)
        if not alive.any():
            break
    if noise_sigma > 0:
        ranges = ranges + np.random.RandomState(seed).normal(0, noise_sigma, num_beams)
    ranges = np.where(ranges >= sensor_max, np.inf, ranges)
    return ranges.astype(np.float32), angles
