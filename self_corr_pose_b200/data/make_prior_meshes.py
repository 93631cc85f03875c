"""Generates self_corr_pose_b200/data/prior_meshes.npz from the reference's category prior OBJ files
(config/<cat>_wild6d/<cat>.obj) -- run once in the build container where /root/reference is mounted.
The priors are benchmark INPUT DATA; they are parsed with a plain
OBJ reader (the files are duplicate-/degenerate-free, so trimesh.load_mesh(process=True) as used by
model/module/mesh.py:66-71 returns the same arrays) and stored un-normalised.
"""
import os
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'


def read_obj(path):
    v, f = [], []
    for line in open(path):
        t = line.split()
        if not t:
            continue
        if t[0] == 'v':
            v.append([float(x) for x in t[1:4]])
        elif t[0] == 'f':
            f.append([int(x.split('/')[0]) - 1 for x in t[1:4]])
    return np.asarray(v, np.float32), np.asarray(f, np.int32)


out = {}
for cat in ['laptop', 'bottle', 'bowl', 'camera', 'mug']:
    v, f = read_obj(os.path.join(REF, 'config', cat + '_wild6d', cat + '.obj'))
    out[cat + '_v'] = v
    out[cat + '_f'] = f
    print(cat, v.shape, f.shape)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'prior_meshes.npz'), **out)
