import mpmath as mp
mp.mp.dps = 60
# f(z) = (asin(sqrt z) - sqrt z) / (sqrt z)^3 on z in [0, zmax]; minimax-ish via Chebyshev interpolation, then check in double
zmax = mp.mpf('0.25')
def f(z):
    if z == 0: return mp.mpf(1)/6
    s = mp.sqrt(z)
    return (mp.asin(s) - s) / (s*z)
import numpy as np
for deg in (9,10,11,12):
    N = deg + 1
    # Chebyshev nodes on [0, zmax]
    nodes = [ (zmax/2)*(1+mp.cos(mp.pi*(2*k+1)/(2*N))) for k in range(N)]
    vals = [f(z) for z in nodes]
    # solve Vandermonde in mp
    A = mp.matrix(N, N)
    for i,z in enumerate(nodes):
        for j in range(N): A[i,j] = z**j
    c = mp.lu_solve(A, mp.matrix(vals))
    cd = [float(x) for x in c]
    # evaluate error of full asin in double arithmetic emulation
    worst = 0
    for x in np.linspace(1e-9, 0.5, 20001):
        z = x*x
        p = 0.0
        for a in reversed(cd): p = p*z + a     # Horner (fma in the kernel: slightly better)
        r = x + x*z*p
        t = mp.asin(mp.mpf(float(x)))
        err = abs((mp.mpf(r) - t)/t)
        worst = max(worst, err)
    print(deg, 'max rel err asin', mp.nstr(worst, 3), 'ulps', mp.nstr(worst/mp.mpf(2)**-53,3))
    if deg == 11:
        print([repr(x) for x in cd])
