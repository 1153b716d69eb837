"""Instrument design helpers (reference marxs/design): Rowland geometry and facet placement."""
from .rowland import (RowlandTorus, ElementsOnTorus, GratingArrayStructure, RectangularGrid,  # noqa: F401
                      CircularMeshGrid, design_tilted_torus)
