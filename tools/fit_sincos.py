"""Coefficients of the fast build's sin / cos on [-pi/4, pi/4] (csrc/mxb_device.cuh: sincos_turn): interpolants at the
Chebyshev nodes in z = x^2, fitted at 60 digits; prints the coefficient tables and the measured error in ulp."""
import mpmath as mp
import numpy as np
mp.mp.dps = 60
zmax = (mp.pi / 4) ** 2 * mp.mpf('1.02')


def fit(f, deg):
    N = deg + 1
    nodes = [(zmax / 2) * (1 + mp.cos(mp.pi * (2 * k + 1) / (2 * N))) for k in range(N)]
    A = mp.matrix(N, N)
    for i, z in enumerate(nodes):
        for j in range(N):
            A[i, j] = z ** j
    c = mp.lu_solve(A, mp.matrix([f(z) for z in nodes]))
    return [float(x) for x in c]


def fs(z):      # (sin(x) - x) / x^3
    if z == 0:
        return -mp.mpf(1) / 6
    x = mp.sqrt(z)
    return (mp.sin(x) - x) / (x * z)


def fc(z):      # (cos(x) - 1 + z/2) / z^2
    if z == 0:
        return mp.mpf(1) / 24
    x = mp.sqrt(z)
    return (mp.cos(x) - 1 + z / 2) / (z * z)


S, C = fit(fs, 5), fit(fc, 5)
ws = wc = 0
for x in np.linspace(-np.pi / 4, np.pi / 4, 40001):
    z = x * x
    p = 0.0
    for a in reversed(S):
        p = p * z + a
    s = x + x * z * p
    p = 0.0
    for a in reversed(C):
        p = p * z + a
    c = (1.0 - 0.5 * z) + z * z * p
    X = mp.mpf(float(x))
    if x != 0:
        ws = max(ws, abs((mp.mpf(s) - mp.sin(X)) / mp.sin(X)))
    wc = max(wc, abs((mp.mpf(c) - mp.cos(X)) / mp.cos(X)))
print('sin', [repr(v) for v in S], 'ulp', mp.nstr(ws * 2 ** 53, 3))
print('cos', [repr(v) for v in C], 'ulp', mp.nstr(wc * 2 ** 53, 3))
