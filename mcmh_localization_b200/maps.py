"""Map loading for the localizer: node:124-177 load_map (host-side setup).

The occupancy grid keeps the ROS OccupancyGrid layout (row = y, index my*W+mx, no flip,
node:136-150); the distance map is the Euclidean distance transform of the free cells in metres
(node:153-157), computed once on the host with SciPy exactly like the node does.
"""
import os
from dataclasses import dataclass

import numpy as np


@dataclass
class GridMap:
    occ: np.ndarray        # (H, W) int8: 0 free, 100 occupied, -1 unknown
    dist: np.ndarray       # (H, W) float32 metres to the nearest non-free cell
    resolution: float
    origin_x: float
    origin_y: float

    @property
    def width(self):
        return int(self.occ.shape[1])

    @property
    def height(self):
        return int(self.occ.shape[0])

    @property
    def limits(self):      # node:168-173
        return np.array([self.origin_x, self.origin_x + self.width * self.resolution,
                         self.origin_y, self.origin_y + self.height * self.resolution])


def map_from_occupancy(occ, resolution, origin_x, origin_y):
    from scipy.ndimage import distance_transform_edt
    occ = np.ascontiguousarray(occ, dtype=np.int8)
    occupancy_binary = (occ != 0).astype(np.uint8)                              # node:153
    dist = (distance_transform_edt(occupancy_binary == 0) * resolution).astype(np.float32)
    return GridMap(occ, np.ascontiguousarray(dist), float(resolution), float(origin_x), float(origin_y))


def read_pgm(path):
    with open(path, "rb") as f:
        data = f.read()
    fields, pos = [], 0
    while len(fields) < 4:
        while data[pos:pos + 1].isspace():
            pos += 1
        if data[pos:pos + 1] == b"#":
            pos = data.index(b"\n", pos) + 1
            continue
        end = pos
        while not data[end:end + 1].isspace():
            end += 1
        fields.append(data[pos:end])
        pos = end
    if fields[0] != b"P5":
        raise ValueError("%s: only binary PGM (P5) is supported" % path)
    w, h, maxval = int(fields[1]), int(fields[2]), int(fields[3])
    dtype = np.uint8 if maxval < 256 else np.dtype(">u2")
    return np.frombuffer(data, dtype=dtype, count=w * h, offset=pos + 1).reshape(h, w)


def load_map_yaml(yaml_path):
    """map_server trinary semantics: occ = (255 - px)/255 (negate: px/255); > occupied_thresh -> 100,
    < free_thresh -> 0, else -1; rows flipped so that cell (0,0) is the bottom-left pixel."""
    import yaml
    with open(yaml_path) as f:
        meta = yaml.safe_load(f)
    img = read_pgm(os.path.join(os.path.dirname(os.path.abspath(yaml_path)), meta["image"]))
    px = img.astype(np.float64)
    p = px / 255.0 if int(meta.get("negate", 0)) else (255.0 - px) / 255.0
    occ = np.full(img.shape, -1, np.int8)
    occ[p > float(meta["occupied_thresh"])] = 100
    occ[p < float(meta["free_thresh"])] = 0
    occ = np.ascontiguousarray(occ[::-1, :])
    return map_from_occupancy(occ, float(meta["resolution"]), float(meta["origin"][0]), float(meta["origin"][1]))


def load_npz(path):
    """Fixture format of tests/golden/map_*.npz (occ, resolution, origin)."""
    g = np.load(path)
    return map_from_occupancy(g["occ"], float(g["resolution"]), float(g["origin"][0]), float(g["origin"][1]))


def tiled_map(base, tiles_x, tiles_y, width=None, height=None):
    """BASELINE config 5: tile a map and crop, keep resolution, centre the origin."""
    occ = np.tile(base.occ, (tiles_y, tiles_x))
    if height:
        occ = occ[:height]
    if width:
        occ = occ[:, :width]
    h, w = occ.shape
    return map_from_occupancy(occ, base.resolution, -0.5 * w * base.resolution, -0.5 * h * base.resolution)
