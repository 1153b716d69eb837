"""Photon sources and pointing (reference marxs/source): photons are BORN on the device.

``source.generate_photons(exposuretime)`` and ``pointing(photons)`` keep the reference's call
pattern (each is one small launch of the trace engine); ``observe(source, pointing, elements,
exposuretime)`` lowers source + pointing + aperture + instrument into ONE program, so a whole
observation is a single kernel launch that reads nothing and writes the event table.  Callable flux /
energy / polarization specifications (user code) are evaluated once per observation and enter the kernel as
input columns; ``poisson_process(rate)`` makes Poisson arrival times on the device."""
from .source import (Source, PointSource, LabPointSourceCone, FarLabPointSource, PointingModel, FixedPointing,  # noqa: F401
                     JitterPointing, RandomArbitraryPdfTable, observe, SourceSpecificationError, poisson_process)
