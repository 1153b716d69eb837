"""Debug helper: C4 chain at small size, probability statistics per kernel / build."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import marxs_b200 as mb
from marxs_b200 import optics, source, _lib

g = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'mlmirror.npz')))
refl = {'X(mm)': g['ml_x_mm'], 'Peak lambda': g['ml_peak_lambda'], 'Peak': g['ml_peak'], 'FWHM(nm)': g['ml_fwhm']}
polt = {'Photon energy': g['ml_pol_energy_ev'], 'Polarization': g['ml_pol']}
a = 2 ** -0.5
rot1 = np.array([[a, 0, -a], [0, 1, 0], [a, 0, a]])
rot2 = np.array([[0, 0, 1.], [a, a, 0], [-a, a, 0]])


def elements():
    return [optics.MultiLayerMirror(reflFile=refl, testedPolarization=polt, orientation=rot1, zoom=[1, 24.5, 12.]),
            optics.FlatBrewsterMirror(orientation=rot2, position=[0., 0., 30.], zoom=[1, 10., 30.]),
            optics.FlatDetector(pixsize=0.05, position=[0., 50., 30.],
                                orientation=np.array([[0, -1., 0], [1., 0, 0], [0, 0, 1.]]), zoom=[1, 20., 20.])]


n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
for jit in ('on', 'off'):
    _lib.load().mxb_set_jit(1 if jit == 'on' else 0)
    for nel in (1, 2, 3):
        srcobj = source.LabPointSourceCone(position=[200., 0, 0], direction=[-1., 0, 0], half_opening=0.02, flux=float(n), energy=0.31)
        mb.set_seed(5)
        o = source.observe(srcobj, None, elements()[:nel], 1., device='cuda', check=False)
        p = o['probability']
        print('jit', jit, 'elements', nel, 'n', len(o), 'prob mean %.6e max %.6e nonzero %d nan %d' % (
            float(p.mean()), float(p.max()), int((p > 0).sum()), int(torch.isnan(p).sum())), _lib.load().mxb_jit_info().decode()[:60])

# the bench's flow: one table reused by successive observations
_lib.load().mxb_set_jit(-1)
srcobj = source.LabPointSourceCone(position=[200., 0, 0], direction=[-1., 0, 0], half_opening=0.02, flux=float(n), energy=0.31)
holder = None
for k in range(3):
    holder = source.observe(srcobj, None, elements(), 1., device='cuda', check=False, out=holder)
    p = holder['probability']
    det = torch.isfinite(holder['det_x'])
    print('reuse call', k, 'prob mean %.6e masked mean %.6e max %.6e nonzero %d' % (float(p.mean()), float(p[det].mean()), float(p.max()), int((p > 0).sum())),
          _lib.load().mxb_jit_info().decode()[:40])
