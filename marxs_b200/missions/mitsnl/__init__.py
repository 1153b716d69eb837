"""CAT gratings from the MKI Space Nanotechnology Laboratory (reference marxs/missions/mitsnl)."""
from .catgrating import (InterpolateEfficiencyTable, NonParallelCATGrating, QualityFactor, L1, L2Abs,  # noqa: F401
                         L2Diffraction, CATL1L2Stack, catsupportbars, l1transtab, l1_order_selector, l1_dims,
                         l2_dims, qualityfactor, d)
