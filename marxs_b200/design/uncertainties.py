"""Random misalignment matrices for the facets of a Parallel (reference marxs/design/uncertainties.py:8-38)."""
import numpy as np

from ..affines import compose, euler2mat

__all__ = ['generate_facet_uncertainty']


def generate_facet_uncertainty(n, xyz, angle_xyz, trans_offset=0, rot_offset=0):
    """``n`` homogeneous (4, 4) matrices: a translation drawn from N(trans_offset, xyz) [mm] and a rotation
    by static-frame Euler angles about x, y, z drawn from N(rot_offset, angle_xyz) [rad].  The draws come
    from numpy's global generator in the reference's order (all translations, then all rotations), so the
    same ``np.random.seed`` reproduces the reference's misalignments."""
    translation = np.random.normal(size=(n, 3), loc=trans_offset, scale=xyz)
    rotation = np.random.normal(size=(n, 3), loc=rot_offset, scale=angle_xyz)
    return [compose(t, euler2mat(a[0], a[1], a[2], 'sxyz'), np.ones(3)) for t, a in zip(translation, rotation)]
