"""Host-side helpers with the reference's names (/root/reference/utils.py:10-191): table printing, graph-pair JSON
loading, pair lists, directory walk and the point-cloud augmentations used by training.  `from utils import *`
in the reference's scripts also hands out `np`, `json`, `math`, `os`, `random`; they are module globals here too.
texttable / matplotlib are optional in this image, so they are imported lazily.
"""
import json
import math
import os
import random

import numpy as np


def tab_printer(args):
    """Print the argument bag as a two-column table (utils.py:10-19)."""
    items = sorted(vars(args).items())
    rows = [["Parameter", "Value"]] + [[name.replace("_", " ").capitalize(), value] for name, value in items]
    try:
        from texttable import Texttable
        table = Texttable()
        table.add_rows(rows)
        print(table.draw())
    except ImportError:
        width = max(len(str(r[0])) for r in rows)
        line = "+" + "-" * (width + 2) + "+" + "-" * 42 + "+"
        print(line)
        for r in rows:
            print(f"| {str(r[0]):<{width}} | {str(r[1])[:40]:<40} |")
            print(line)


def _read_graph(path):
    with open(path) as handle:
        return json.load(handle)


def process_pair(path, cache=None):
    """Two graph JSON files -> one pair dict; `distance` is the planar pose distance (utils.py:21-38: pose[3], pose[11]).
    `cache` (a dict, optional): parsed files are kept and reused — the reference re-reads both files for every listed
    pair of every epoch; the returned dict only references the parsed lists, nothing downstream mutates them."""
    if cache is None:
        first, second = _read_graph(path[0]), _read_graph(path[1])
    else:
        first = cache.get(path[0])
        if first is None:
            first = cache[path[0]] = _read_graph(path[0])
        second = cache.get(path[1])
        if second is None:
            second = cache[path[1]] = _read_graph(path[1])
    dx = first["pose"][3] - second["pose"][3]
    dz = first["pose"][11] - second["pose"][11]
    return {"centers_1": first["centers"], "nodes_1": first["nodes"],
            "centers_2": second["centers"], "nodes_2": second["nodes"],
            "distance": math.sqrt(dx ** 2 + dz ** 2)}


def load_paires(file, graph_pairs_dir):
    """Pair list file ("a.json b.json" per line) -> [[dir/a.json, dir/b.json], ...] (utils.py:61-70)."""
    pairs = []
    with open(file) as handle:
        for line in handle:
            if not line:
                break
            names = line.strip().split(" ")
            pairs.append([os.path.join(graph_pairs_dir, names[0]), os.path.join(graph_pairs_dir, names[1])])
    return pairs


def listDir(path, list_name):
    """Append every file under `path` (recursively) to list_name (utils.py:73-84)."""
    for entry in os.listdir(path):
        full = os.path.join(path, entry)
        if os.path.isdir(full):
            listDir(full, list_name)
        else:
            list_name.append(full)


# ---- augmentations (utils.py:86-178).  Same distributions and the same order of RNG draws as the reference. ----

def flip_point_cloud(batch_data):
    if random.random() > 0.5:
        batch_data[:, :, 0] = -batch_data[:, :, 0]
    return batch_data


def _rot_z(angle):
    c, s = np.cos(angle), np.sin(angle)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def rotate_point_cloud(batch_data):
    """Random rotation about the up (z) axis, one angle per cloud; returns float32 like the reference."""
    out = np.zeros(batch_data.shape, dtype=np.float32)
    for b in range(batch_data.shape[0]):
        angle = np.random.uniform() * 2 * np.pi
        out[b, ...] = np.dot(batch_data[b, ...].reshape((-1, 3)), _rot_z(angle))
    return out


def jitter_point_cloud(batch_data, sigma=0.01, clip=0.05):
    assert clip > 0
    b, n, c = batch_data.shape
    noise = np.clip(sigma * np.random.randn(b, n, c), -1 * clip, clip)
    noise += batch_data
    return noise


def random_scale_point_cloud(batch_data, scale_low=0.8, scale_high=1.25):
    scales = np.random.uniform(scale_low, scale_high, batch_data.shape[0])
    for b in range(batch_data.shape[0]):
        batch_data[b, :, :] *= scales[b]
    return batch_data


def rotate_perturbation_point_cloud(batch_data, angle_sigma=0.015, angle_clip=0.045):
    out = np.zeros(batch_data.shape, dtype=np.float32)
    for b in range(batch_data.shape[0]):
        ax, ay, az = np.clip(angle_sigma * np.random.randn(3), -angle_clip, angle_clip)
        rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
        ry = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
        rz = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]])
        out[b, ...] = np.dot(batch_data[b, ...].reshape((-1, 3)), np.dot(rz, np.dot(ry, rx)))
    return out


def shift_point_cloud(batch_data, shift_range=0.3):
    shifts = np.random.uniform(-shift_range, shift_range, (batch_data.shape[0], 3))
    for b in range(batch_data.shape[0]):
        batch_data[b, :, :] += shifts[b, :]
    return batch_data


def vis_point_cloud(pc):
    import matplotlib.pyplot as plt
    from mpl_toolkits.mplot3d import Axes3D  # noqa: F401
    pts = pc[0, :, :]
    fig = plt.figure()
    ax = fig.add_subplot(projection="3d")
    ax.scatter(pts[:, 0], pts[:, 1], pts[:, 2], c="b", marker=".", s=10, linewidth=0, alpha=1)
    plt.show()
