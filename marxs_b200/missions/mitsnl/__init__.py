"""CAT gratings from the MKI Space Nanotechnology Laboratory (reference marxs/missions/mitsnl)."""
from .catgrating import InterpolateEfficiencyTable, NonParallelCATGrating
