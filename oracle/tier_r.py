"""Tier-R oracle loader: import the UNMODIFIED reference under the stand-ins.

TEST INFRASTRUCTURE.  Only ``oracle/gen_golden.py`` and the oracle
cross-check tests use this, and only in the build container:
``/root/reference`` does not exist on the GPU box, so nothing in ``-m gpu``
tests, ``smoke()`` or ``bench.py`` may call it (``available()`` is False there).
"""
import importlib
import importlib.metadata
import os
import sys

REFERENCE_ROOT = os.environ.get('MARXS_REFERENCE_ROOT', '/root/reference')
_STANDIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'standin')


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'marxs'))


def load_reference():
    """Return the imported reference package ``marxs`` (with sub-packages
    math, optics, simulator, missions.chandra importable)."""
    if not available():
        raise RuntimeError('reference tree not present at ' + REFERENCE_ROOT)
    if 'marxs' in sys.modules and getattr(sys.modules['marxs'], '_tier_r', False):
        return sys.modules['marxs']
    try:
        import astropy  # noqa: F401  (a real install wins if there is one)
        import transforms3d  # noqa: F401
    except ImportError:
        for m in list(sys.modules):
            if m.split('.')[0] in ('astropy', 'transforms3d'):
                del sys.modules[m]
        if _STANDIN not in sys.path:
            sys.path.insert(0, _STANDIN)
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)
    # marxs/__init__.py:4-8 swallows PackageNotFoundError and then
    # marxs/base/base.py:13 fails on the missing __version__ for a source tree.
    real_version = importlib.metadata.version

    def version(name):
        if name == 'marxs':
            return '2.0.dev0'
        return real_version(name)

    importlib.metadata.version = version
    try:
        marxs = importlib.import_module('marxs')
        for sub in ('marxs.base', 'marxs.math.geometry', 'marxs.math.polarization',
                    'marxs.optics', 'marxs.simulator'):
            importlib.import_module(sub)
    finally:
        importlib.metadata.version = real_version
    marxs._tier_r = True
    return marxs
