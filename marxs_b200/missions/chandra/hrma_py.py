"""Pure-engine approximation of the Chandra mirror (reference marxs/missions/chandra/hrma_py.py)."""
import numpy as np

from ... import optics
from .data import NOMINAL_FOCALLENGTH

__all__ = ['mirror_radii', 'Aperture', 'HRMA']

mirror_radii = np.array([[598., 610.], [481, 491], [424, 433], [315, 322]])


class Aperture(optics.MultiAperture):
    """Four ring openings above the mirror shells (reference :16-26)."""

    def __init__(self, **kwargs):
        if 'id_col' not in kwargs:
            kwargs['id_col'] = 'mirror_shell'
        if 'elements' not in kwargs:
            kwargs['elements'] = [optics.CircleAperture(position=[NOMINAL_FOCALLENGTH, 0, 0],
                                                        zoom=[1, r[1], r[1]], r_inner=r[0])
                                  for r in mirror_radii]
        super().__init__(**kwargs)


class HRMA(optics.FlatStack):
    """PerfectLens + RadialMirrorScatter + a constant 0.66 throughput (reference :29-50)."""

    def __init__(self, **kwargs):
        kwargs['position'] = [NOMINAL_FOCALLENGTH, 0, 0]
        kwargs['zoom'] = [1, 650., 650]
        kwargs['elements'] = [optics.PerfectLens, optics.RadialMirrorScatter, optics.EnergyFilter]
        kwargs['keywords'] = [{'focallength': NOMINAL_FOCALLENGTH},
                              {'inplanescatter': 3.6e-6, 'perpplanescatter': 1.2e-6},
                              {'filterfunc': 0.66, 'name': 'support spider, reflectivity, et al.'}]
        super().__init__(**kwargs)
