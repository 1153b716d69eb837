"""Debug helper: culled vs exhaustive search on the random array of tests/test_parity_round2.py (seed argument)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import marxs_b200 as mb
from marxs_b200 import optics, simulator, program, _lib
from oracle import marxs_oracle as mo
from test_parity_gpu import make_photons, SEED

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 4
rng = np.random.default_rng(SEED + 180 + seed)
n = 40000
layers = int(rng.integers(1, 3))
pitch = float(rng.choice([8.5, 10., 14.]))
half = float(rng.uniform(3.5, 0.62 * pitch))
pos4ds = []
for layer in range(layers):
    for y in np.arange(-40, 41, pitch):
        for z in np.arange(-30, 31, pitch):
            p = [-7. * layer + 0.003 * rng.uniform(0, 1) * (y * y + z * z), y + 0.4 * pitch * layer + rng.uniform(-1, 1),
                 z - 0.3 * pitch * layer + rng.uniform(-1, 1)]
            a, b = rng.uniform(-0.12, 0.12, 2)
            Ry = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
            Rz = np.array([[np.cos(b), -np.sin(b), 0], [np.sin(b), np.cos(b), 0], [0, 0, 1.]])
            pos4ds.append(mo.compose(p, Rz @ Ry, [1., half * rng.uniform(0.8, 1.), half * rng.uniform(0.8, 1.)]))
order = rng.permutation(len(pos4ds))
pos4ds = [pos4ds[k] for k in order]
print('seed', seed, 'layers', layers, 'pitch', pitch, 'half', half, 'F', len(pos4ds))
if seed % 2:
    sel = optics.OrderSelector(orderlist=np.array([-30, -12, 0, 9, 25]), p=np.full(5, 0.2))
    make = lambda: simulator.Parallel(elem_class=optics.FlatGrating, elem_pos=pos4ds, id_col='facet', elem_args=dict(d=2e-4, order_selector=sel))
else:
    make = lambda: simulator.Parallel(elem_class=optics.FlatDetector, elem_pos=pos4ds, id_col='facet', elem_args=dict(pixsize=0.05))
table = make_photons(rng, n, spread=float(rng.choice([0.05, 0.3, 1.2])), x0=40., lateral=45., e_lo=0.8, e_hi=1.6)
G = np.array([program.geom14(p) for p in pos4ds])
g = program.build_cull_grid(G)
cert = program.single_hit_successors(G, g)
print('T2', g['T2'], 'cert t2', cert[0], 'lists', None if cert[2] is None else (len(cert[2]), int(np.diff(cert[1]).max())))
_lib.load().mxb_set_jit(2)
res = []
for ex in (False, True):
    program.EXHAUSTIVE_SEARCH = ex
    mb.set_seed(5 + seed)
    res.append(make()(mb.PhotonBatch(table, device='cuda')).to_numpy())
program.EXHAUSTIVE_SEARCH = False
a, b = res
bad = np.nonzero(a['facet'] != b['facet'])[0]
print('facet mismatches', len(bad), 'of', n)
d = table['dir'][:, :3]
nb = g['nbar']
dn = d @ nb
tan = np.sqrt(np.maximum((d * d).sum(1) - dn * dn, 0)) / np.abs(dn)
# oracle: sequential loop over all facets
cur = table.copy()
orac = mo.Parallel(mo.FlatDetector, pos4ds, dict(pixsize=0.05), id_col='facet') if seed % 2 == 0 else None
if orac is not None:
    want = orac(cur)
    print('culled vs oracle', (a['facet'] != want['facet']).sum(), ' exhaustive vs oracle', (b['facet'] != want['facet']).sum())
for i in bad[:12]:
    hits = [j for j in range(len(pos4ds)) if mo.plane_intersect(mo.PlaneConsts(pos4ds[j]), table['dir'][i:i + 1], table['pos'][i:i + 1])[0][0]]
    print(i, 'culled', a['facet'][i], 'exhaustive', b['facet'][i], 'tan %.3f' % tan[i], 'facets on the initial ray', hits,
          'successors', None if cert[2] is None else [list(cert[2][cert[1][h]:cert[1][h + 1]]) for h in hits])
for c in b:
    same = np.array_equal(a[c], b[c], equal_nan=True)
    if not same:
        x, y = np.asarray(a[c], dtype=float), np.asarray(b[c], dtype=float)
        if x.ndim > 1:
            diff = np.nonzero(~((x == y) | (np.isnan(x) & np.isnan(y))).all(axis=1))[0]
        else:
            diff = np.nonzero(~((x == y) | (np.isnan(x) & np.isnan(y))))[0]
        print('column', c, 'differs at', len(diff), 'photons; first', diff[:5], x[diff[:3]], y[diff[:3]], 'facet', a['facet'][diff[:5]], 'tan', tan[diff[:5]])
