"""Geometry layer: ``rotate_mat`` / ``rotateXYZ_mat`` with the reference's signatures
(pynmap/rotation.py:14-119).  The matrices are built on the host with the reference's
own numpy expressions (so float32-vs-float64 trigonometry behaves identically); the
application ``mat . pos^T`` runs in the pnm_rotate kernel with the rounding of np.dot.
"""
import numpy as np
import torch

from . import engine


def rotation_matrix(PA=0.0, inclination=0.0, direct=True):
    """float32 3x3 of rotate_mat (rotation.py:99-113).  Angles in degrees; when they arrive
    as np.float32 (snapshot.rotate does that, misc_io.py:363) the trigonometry is float32."""
    PA = np.deg2rad(PA)
    inclination = np.deg2rad(inclination)
    cPA, sPA = np.cos(PA), np.sin(PA)
    ci, si = np.cos(inclination), np.sin(inclination)
    mat = np.array([[cPA, -sPA, 0], [ci * sPA, ci * cPA, -si], [si * sPA, cPA * si, ci]], dtype=np.float32)
    return mat if direct else np.transpose(mat)


def rotation_matrix_xyz(alpha=0.0, beta=0.0, gamma=0.0, rotorder=[0, 1, 2], direct=True):
    """float32 3x3 of rotateXYZ_mat (rotation.py:32-65): float32 product of the axis matrices."""
    ar, br, gr = np.deg2rad(alpha), np.deg2rad(beta), np.deg2rad(gamma)
    ca, sa, cb, sb, cg, sg = np.cos(ar), np.sin(ar), np.cos(br), np.sin(br), np.cos(gr), np.sin(gr)
    mats = [np.array([[1, 0, 0], [0, ca, sa], [0, -sa, ca]], dtype=np.float32),
            np.array([[cb, 0, -sb], [0, 1, 0], [sb, 0, cb]], dtype=np.float32),
            np.array([[cg, sg, 0], [-sg, cg, 0], [0, 0, 1]], dtype=np.float32)]
    if not direct:
        mats = [np.transpose(m) for m in mats]
    return np.dot(np.dot(mats[rotorder[0]], mats[rotorder[1]]), mats[rotorder[2]])


def _soa_on_device(a, dtype):
    """(N, 3) numpy/torch -> (3, N) contiguous device tensor.  Column-contiguous input
    (what the reference itself produces, rotation.py:115) is taken without a transpose copy."""
    if isinstance(a, torch.Tensor):
        t = a.detach()
    else:
        t = torch.from_numpy(np.asarray(a))
    if t.dim() != 2 or t.shape[1] != 3:
        raise ValueError("shapes %s and (3, 3) not aligned: expected an (N, 3) array" % (tuple(t.shape),))
    return t.to(device=engine.current_device(), dtype=dtype).t().contiguous()


def apply_matrix(mat, pos, vel=None):
    """(mat . pos^T)^T and (mat . vel^T)^T on the device; returns arrays of the input kind
    (numpy in -> numpy out, CUDA tensor in -> CUDA tensor out), laid out column-contiguous
    like the reference's results."""
    want_torch = engine.is_cuda_tensor(pos)
    outs = []
    for a in (pos, vel):
        if a is None:
            outs.append(None)
            continue
        dtype = engine.result_dtype(a)
        soa = _soa_on_device(a, dtype)
        out = engine.rotate_soa((soa[0], soa[1], soa[2]), mat)
        outs.append(out.t() if want_torch else out.cpu().numpy().T)
    return outs[0], outs[1]


def rotate_mat(pos, vel=None, PA=0.0, inclination=0.0, direct=True):
    """Rotation of PA (around z) and inclination (around x), angles in degrees
    (pynmap/rotation.py:73-119).  Returns (pos, vel); vel may be None."""
    return apply_matrix(rotation_matrix(PA, inclination, direct), pos, vel)


def rotateXYZ_mat(pos, vel=None, alpha=0.0, beta=0.0, gamma=0.0, rotorder=[0, 1, 2], direct=True):
    """Rotation around the three axes in ``rotorder`` (pynmap/rotation.py:14-71)."""
    return apply_matrix(rotation_matrix_xyz(alpha, beta, gamma, rotorder, direct), pos, vel)
