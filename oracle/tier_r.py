"""Tier-R oracle loader: import the UNMODIFIED reference under the stand-ins.

TEST INFRASTRUCTURE.  Two places hold the reference: ``/root/reference`` (the source tree, build
container only: ``oracle/gen_golden.py`` and the oracle cross-check tests read it) and
``baseline/_ref`` (the offline ``pip install --target`` of the same tree made by
``baseline/install_ref.py``; git-ignored, it travels to the GPU box with the snapshot).  The GPU box
has no ``/root/reference``: there only ``bench.py``'s CPU legs (``--impl reference`` and
``cpu_baseline``) import the installed copy, as the checker / baseline, never on the product path.
"""
import importlib
import importlib.metadata
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
INSTALLED_ROOT = os.path.join(os.path.dirname(_HERE), 'baseline', '_ref')
_STANDIN = os.path.join(_HERE, 'standin')


def _pick_root():
    env = os.environ.get('MARXS_REFERENCE_ROOT')
    for c in ([env] if env else []) + ['/root/reference', INSTALLED_ROOT]:
        if os.path.isdir(os.path.join(c, 'marxs')):
            return c
    return env or '/root/reference'


REFERENCE_ROOT = _pick_root()


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'marxs'))


def installed_available():
    """The pip-installed copy under baseline/_ref (the one that exists on the GPU box)."""
    return os.path.isdir(os.path.join(INSTALLED_ROOT, 'marxs'))


def load_reference(root=None):
    """Return the imported reference package ``marxs`` (with sub-packages
    math, optics, simulator, missions.chandra importable).  ``root``: directory that holds the
    ``marxs`` package (default: the source tree if present, else the installed copy)."""
    global REFERENCE_ROOT
    if root is not None:
        REFERENCE_ROOT = root
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
