"""Instrument design helpers (reference marxs/design): Rowland geometry, facet placement, tolerancing drivers."""
from .rowland import (RowlandTorus, ElementsOnTorus, GratingArrayStructure, RectangularGrid,  # noqa: F401
                      CircularMeshGrid, design_tilted_torus)
from .uncertainties import generate_facet_uncertainty  # noqa: F401
from . import tolerancing  # noqa: F401
