#!/usr/bin/env python
"""Install the UNMODIFIED reference into baseline/_ref (git-ignored; ships to the GPU box with the snapshot).

    python baseline/install_ref.py

Runs the contract's offline recipe

    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
        --target baseline/_ref <copy of /root/reference>

from a copy under /tmp (the build writes egg-info into the source tree and /root/reference is read-only)
with --no-deps (astropy / transforms3d / x3d are not in the wheelhouse; the CPU arm imports the
installed package under the stand-ins of oracle/standin, see oracle/tier_r.py).

The reference selects its package data (HESSdesign.rdb, the multilayer tables, the CAT efficiency tables ...)
through setuptools_scm's git file finder; /root/reference has no .git, so the wheel built here contains the
Python modules only.  The second step completes the install with exactly those non-Python package files,
which a build from a VCS checkout ships.  Nothing below baseline/_ref is tracked by git or imported by the
product (marxs_b200/); it is the reference arm of bench.py and an extra pin of the oracle."""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, '_ref')
SOURCE = os.environ.get('MARXS_REFERENCE_ROOT', '/root/reference')
SKIP_EXT = ('.py', '.pyc', '.pyo', '.c', '.o', '.so')


def install():
    if not os.path.isdir(os.path.join(SOURCE, 'marxs')):
        print('reference tree not present at {0}: keeping the existing baseline/_ref'.format(SOURCE))
        return os.path.isdir(os.path.join(TARGET, 'marxs'))
    tmp = tempfile.mkdtemp(prefix='marxs_ref_')
    try:
        src = os.path.join(tmp, 'reference')
        shutil.copytree(SOURCE, src)
        shutil.rmtree(TARGET, ignore_errors=True)
        cmd = [sys.executable, '-m', 'pip', 'install', '--quiet', '--no-index', '--no-build-isolation', '--no-deps',
               '--find-links', '/opt/wheelhouse', '--target', TARGET, src]
        subprocess.check_call(cmd)
        n = 0
        pkg = os.path.join(SOURCE, 'marxs')
        for root, dirs, files in os.walk(pkg):
            dirs[:] = [d for d in dirs if d != '__pycache__']
            for f in files:
                if f.endswith(SKIP_EXT):
                    continue
                rel = os.path.relpath(os.path.join(root, f), pkg)
                dst = os.path.join(TARGET, 'marxs', rel)
                if not os.path.exists(dst):
                    os.makedirs(os.path.dirname(dst), exist_ok=True)
                    shutil.copyfile(os.path.join(root, f), dst)
                    n += 1
        print('installed the reference into {0} (pip --no-deps) + {1} package-data files'.format(TARGET, n))
        return True
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == '__main__':
    sys.exit(0 if install() else 1)
