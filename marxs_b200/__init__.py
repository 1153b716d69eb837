"""marxs_b200 — B200-native photon ray-trace engine behind the MARXS optical-element API."""
__version__ = '0.1.0'

energy2wave = 1.2398419292004202e-06
'''Convert from photon energy in keV to wavelength in mm (reference marxs/__init__.py:14).'''

from .photons import PhotonBatch, generate_test_photons  # noqa: E402,F401
from .rng import set_seed, inject_draws  # noqa: E402,F401
