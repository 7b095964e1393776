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


def face_centres(faces, pts):
    return np.array([pts[q].mean(0) for q in faces])


def min_edge_length(faces, pts):
    """gMin(minEdgeLengths()) of directMapInterfaceToInterfaceMapping.C:560-597: every point's shortest edge, then the
    minimum over the points = the shortest edge of the patch."""
    return min(np.linalg.norm(pts[q[i]] - pts[q[(i + 1) % len(q)]]) for q in faces for i in range(len(q)))


def direct_map_cases():
    """name -> (to, from, tol, expected map or None)"""
    rng = np.random.default_rng(42)
    cases = {}
    faces, pts = grid_patch(np.linspace(0, 2, 41), np.linspace(0, 1, 26))          # 1000 faces: several tiles of 256
    cA = face_centres(faces, pts)
    tol = 0.001 * min_edge_length(faces, pts)
    perm = rng.permutation(len(faces))
    cB = cA[perm] + 0.3 * tol * rng.standard_normal((len(faces), 3)) / np.sqrt(3)
    cases["permuted_faces"] = (cB, cA, tol, perm.astype(np.int32))
    cases["permuted_points"] = (pts[rng.permutation(len(pts))], pts, tol, None)
    # distances straddling the tolerance, exactly representable: tol = 2^-10
    t = 2.0 ** -10
    frm = np.array([[0.0, 0, 0], [1.0, 0, 0], [2.0, 0, 0]])
    to = np.array([[t, 0, 0], [np.nextafter(t, 0), 0, 0], [1.0, np.nextafter(t, 1), 0], [2.0, 0, -0.5 * t], [5.0, 5, 5]])
    cases["at_the_tolerance"] = (to, frm, t, np.array([-1, 0, -1, 2, -1], np.int32))
    # several candidates within the tolerance: the first one wins, even when a later one is closer
    frm = np.concatenate([rng.random((300, 3)) + 10.0, [[0.4, 0, 0], [0.1, 0, 0], [0.0, 0, 0]], rng.random((300, 3)) + 10.0])
    cases["first_match_wins"] = (np.zeros((1, 3)), frm, 0.5, np.array([300], np.int32))
    cases["empty_from"] = (cA[:5], np.zeros((0, 3)), tol, np.full(5, -1, np.int32))
    cases["empty_to"] = (np.zeros((0, 3)), cA, tol, np.zeros(0, np.int32))
    cases["zero_tol"] = (cA[:7], cA[:7], 0.0, np.full(7, -1, np.int32))        # mag(0) < 0 is false
    return cases
