"""Patch generators for the GGI weight tests: faces as (offsets, labels) + points, master normals +z / outward, slave
normals opposite (the two sides of an interface face each other)."""
import numpy as np


def grid_patch(x, y, z=0.0, flip=False, warp=None):
    """Quads over the tensor grid x (i-fastest) x y; flip: reversed point order (normal -z)."""
    x, y = np.asarray(x, float), np.asarray(y, float)
    X, Y = np.meshgrid(x, y)
    Z = np.full_like(X, z) if warp is None else z + warp(X, Y)
    pts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1)
    nx = x.size
    faces = []
    for j in range(y.size - 1):
        for i in range(x.size - 1):
            q = [j * nx + i, j * nx + i + 1, (j + 1) * nx + i + 1, (j + 1) * nx + i]
            faces.append(q[::-1] if flip else q)
    return faces, pts


def split_triangles(faces):
    out = []
    for q in faces:
        out += [[q[0], q[1], q[2]], [q[0], q[2], q[3]]]
    return out


def to_csr(faces):
    off = np.zeros(len(faces) + 1, np.int32)
    off[1:] = np.cumsum([len(f) for f in faces])
    lab = np.array([p for f in faces for p in f], np.int32)
    return off, lab
