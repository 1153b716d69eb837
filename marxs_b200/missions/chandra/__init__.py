"""Chandra: HRMA (pure-engine approximation), HETG facet array, ACIS chips
(reference marxs/missions/chandra)."""
from .hess import HETG
from .det_acis import ACIS, ACISChip
from .data import NOMINAL_FOCALLENGTH, AIMPOINTS, PIXSIZE
from .hrma_py import Aperture, HRMA
