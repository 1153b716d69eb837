#!/bin/bash
# compute-sanitizer pass over the trace kernel (run on the GPU box):  bash tools/sanitize.sh
set -e
cd "$(dirname "$0")/.."
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool"
  compute-sanitizer --tool $tool --error-exitcode 1 python - <<'PY'
import numpy as np
import __graft_entry__ as g
import marxs_b200 as mb
from marxs_b200 import optics, simulator, host as mhost
g.smoke()
# small arrays in brute-force mode, an aperture, a baffle and the host-buffer path
rng = np.random.default_rng(0)
n = 3000
d = np.zeros((n, 4)); d[:, 0] = -1
src = {'dir': d, 'energy': np.ones(n), 'polarization': np.tile([0., 1, 0, 0], (n, 1)), 'probability': np.ones(n)}
inst = simulator.Sequence(elements=[optics.RectangleAperture(position=[50., 0, 0], zoom=[1, 5, 5]),
                                    optics.Baffle(position=[40., 0, 0], zoom=[1, 4, 4]),
                                    simulator.Parallel(elem_class=optics.FlatDetector, id_col='CCD',
                                                       elem_pos={'position': [[0., -3, 0], [0., 3, 0]]},
                                                       elem_args={'pixsize': 0.1, 'zoom': [1, 2.5, 5]})])
out = inst(mb.PhotonBatch(src, device='cuda')).to_numpy()
assert (out['CCD'] >= 0).sum() > 100
ht = mhost.HostPhotonTable.from_columns({k: out[k] if k in out else v for k, v in src.items()} | {'pos': np.tile([60., 0, 0, 1.], (n, 1))})
mhost.trace_host(optics.FlatDetector(pixsize=0.1, zoom=[1, 5, 5]), ht, chunk=1024)
# specialised kernel of the same small programs, event compaction, CAT stack + cylinder detector
from marxs_b200 import _lib, events
from marxs_b200.missions import mitsnl
_lib.load().mxb_set_jit(2)
out2 = inst(mb.PhotonBatch(src, device='cuda'))
assert _lib.load().mxb_jit_info().startswith(b'jit ')
ev = events.compact(out2, sel='CCD', weight='probability')
assert len(ev) == int(((out2['CCD'] >= 0) & (out2['probability'] > 0)).sum())
sel = optics.OrderSelector(np.arange(-2, 3))
cat = simulator.Sequence(elements=[simulator.Parallel(elem_class=mitsnl.CATL1L2Stack, id_col='facet',
                                                       elem_pos={'position': [[0., -3, 0], [0., 3, 0]]},
                                                       elem_args={'zoom': [1, 2.5, 5], 'order_selector': sel}),
                                    optics.CircularDetector(pixsize=0.1, position=[-50., 0, 0], zoom=[50., 50., 10.])])
ph = {k: v for k, v in src.items()} | {'pos': np.tile([60., 0, 0, 1.], (n, 1)) + rng.uniform(-4, 4, (n, 4)) * [0, 1, 1, 0]}
for jit in (2, 0):
    _lib.load().mxb_set_jit(jit)
    o = cat(mb.PhotonBatch(ph, device='cuda')).to_numpy()
    assert np.isfinite(o['det_phi']).sum() > 100
# lens with a reflectivity table (global-memory table lookups) in both kernels
rt = optics.ReflectivityTable(np.linspace(0.5, 2., 7), np.linspace(0., 0.2, 9), rng.uniform(0.3, 1., (7, 9)))
lens = optics.PerfectLens(focallength=30., zoom=[1, 10, 10], reflectivity_interpolator=rt)
for jit in (2, 0):
    _lib.load().mxb_set_jit(jit)
    o = lens(mb.PhotonBatch(ph, device='cuda')).to_numpy()
    assert (o['probability'] < 1).sum() > 100
_lib.load().mxb_set_jit(-1)
# sigma clipping kernels (radix-select median) and the exported device math
from marxs_b200 import analysis
import torch
x = np.concatenate([rng.normal(0., 1., 20001), rng.uniform(-40, 40, 200), [np.nan, np.inf]])
m = analysis.sigma_clipped_stats(torch.as_tensor(x, device='cuda'))
assert abs(m[1]) < 0.1 and 0.9 < m[2] < 1.1
print('sanitizer workload ok')
PY
done
