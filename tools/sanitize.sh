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
print('sanitizer workload ok')
PY
done
