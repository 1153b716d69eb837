import math
import numpy as np


def compose(T, R, Z, S=None):
    """4x4 affine from translation T, rotation R, zooms Z, shears S."""
    n = len(T)
    R = np.asarray(R)
    A = np.eye(n + 1)
    ZS = np.diag(Z).astype(float)
    if S is not None:
        S = np.asarray(S)
        ZSS = np.eye(n)
        ZSS[np.triu_indices(n, 1)] = S
        ZS = np.dot(ZS, ZSS)
    A[:n, :n] = np.dot(R, ZS)
    A[:n, n] = T
    return A


def decompose44(A44):
    """Gram-Schmidt decomposition of a 4x4 affine into T, R, Z, S."""
    A44 = np.asarray(A44)
    T = A44[:-1, -1]
    RZS = A44[:-1, :-1]
    M0, M1, M2 = np.array(RZS).T
    sx = math.sqrt(np.sum(M0 ** 2))
    M0 /= sx
    sx_sxy = np.dot(M0, M1)
    M1 -= sx_sxy * M0
    sy = math.sqrt(np.sum(M1 ** 2))
    M1 /= sy
    sxy = sx_sxy / sx
    sx_sxz = np.dot(M0, M2)
    sy_syz = np.dot(M1, M2)
    M2 -= (sx_sxz * M0 + sy_syz * M1)
    sz = math.sqrt(np.sum(M2 ** 2))
    M2 /= sz
    sxz = sx_sxz / sx
    syz = sy_syz / sy
    Rmat = np.array([M0, M1, M2]).T
    if np.linalg.det(Rmat) < 0:
        sx *= -1
        Rmat[:, 0] *= -1
    return T, Rmat, np.array([sx, sy, sz]), np.array([sxy, sxz, syz])


def decompose(A):
    A = np.asarray(A)
    T = A[:-1, -1]
    RZS = A[:-1, :-1]
    ZS = np.linalg.cholesky(np.dot(RZS.T, RZS)).T
    Z = np.diag(ZS).copy()
    shears = ZS / Z[:, np.newaxis]
    n = len(Z)
    S = shears[np.triu(np.ones((n, n)), 1).astype(bool)]
    R = np.dot(RZS, np.linalg.inv(ZS))
    if np.linalg.det(R) < 0:
        Z[0] *= -1
        ZS[0] *= -1
        R = np.dot(RZS, np.linalg.inv(ZS))
    return T, R, Z, S
