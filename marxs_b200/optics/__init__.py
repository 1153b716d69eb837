"""Optical elements (same names as ``marxs.optics``)."""
from .aperture import RectangleAperture, CircleAperture, MultiAperture
from .detector import FlatDetector, CircularDetector
from .grating import FlatGrating, CATGrating, OrderSelector, EfficiencyFile
from .mirror import PerfectLens, ReflectivityTable
from .baffles import Baffle, CircularBaffle
from .scatter import RadialMirrorScatter, RandomGaussianScatter
from .filter import EnergyFilter, GlobalEnergyFilter, Tabulated1D
from .base import OpticalElement, FlatOpticalElement, FlatStack, photonlocalcoords
from .multiLayerMirror import FlatBrewsterMirror, MultiLayerEfficiency, MultiLayerMirror
