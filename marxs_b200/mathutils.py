"""Vector helpers of the reference's marxs/math/utils.py (:10-64, :150-164) for user elements and tests.

Every function takes numpy arrays or torch tensors (CPU or CUDA) and returns the same kind: inside a user
``specific_process_photons`` the arguments are device tensors and nothing leaves the device."""
import numpy as np
import torch

__all__ = ['xyz2zxy', 'e2h', 'h2e', 'distance_point_point', 'norm_vector']

xyz2zxy = np.array([[0., 1., 0., 0.],
                    [0., 0., 1., 0.],
                    [1., 0., 0., 0.],
                    [0., 0., 0., 1.]])
'''Flips coordinates for missions whose optical axis is the z axis.'''


def _is_t(x):
    return isinstance(x, torch.Tensor)


def e2h(e, w):
    """Euclidean -> homogeneous coordinates: ``w`` = 0 for directions, 1 for positions (reference :10-35)."""
    if not ((w == 0) or (w == 1)):
        raise ValueError('w must be 0 or 1.')
    if _is_t(e):
        e = e.as_subclass(torch.Tensor)
        return torch.cat([e, torch.full(tuple(e.shape[:-1]) + (1,), float(w), dtype=e.dtype, device=e.device)], dim=-1)
    e = np.asarray(e)
    h = np.empty(e.shape[:-1] + (e.shape[-1] + 1,))
    h[..., :3] = e
    h[..., 3] = w
    return h


def h2e(h):
    """Homogeneous -> Euclidean coordinates (reference :38-64): all points at infinity or all finite points."""
    if _is_t(h):
        h = h.as_subclass(torch.Tensor)
        w = h[..., 3]
        if bool((w == 0).all()) or bool(torch.isclose(w, torch.ones_like(w)).all()):
            return h[..., :3]
        if bool((w != 0).all()):
            return h[..., :3] / w[..., None]
        raise ValueError('Input array must be either all Euclidean points or all points at infinity.')
    h = np.asarray(h)
    if np.all(h[..., 3] == 0) or np.allclose(h[..., 3], 1):
        return h[..., :3]
    if np.all(h[..., 3] != 0):
        return h[..., :3] / h[..., 3][..., None]
    raise ValueError('Input array must be either all Euclidean points or all points at infinity.')


def distance_point_point(h_p1, h_p2):
    """Euclidean distance between points given in homogeneous coordinates, shape (4,) or (N, 4) (reference :67-90)."""
    nd = h_p1.dim() if _is_t(h_p1) else np.ndim(h_p1)
    if nd == 1:
        axis = 0
    elif nd == 2:
        axis = 1
    else:
        raise ValueError('This function expects 1d or 2d input.')
    d = h2e(h_p1) - h2e(h_p2)
    return torch.linalg.norm(d, dim=axis) if _is_t(d) else np.linalg.norm(d, axis=axis)


def norm_vector(vec):
    """Normalise Euclidean vectors of shape (n, 3) (reference :150-164; also used on (n, 4) directions with w = 0)."""
    if _is_t(vec):
        vec = vec.as_subclass(torch.Tensor)
        return vec / torch.sqrt(torch.sum(vec * vec, dim=-1))[:, None]
    vec = np.asarray(vec)
    return vec / np.sqrt(np.sum(vec * vec, axis=-1))[:, None]
