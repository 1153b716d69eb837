"""Minimal stand-in for astropy: containers and constants only.
TEST INFRASTRUCTURE — see ../README.md.  Not a reimplementation of astropy."""
__version__ = '0.0.standin'
from . import units, table, config, utils, coordinates, time, io  # noqa: F401
