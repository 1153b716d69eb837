import inspect
import os
import sys


def get_pkg_data_filename(data_name, package=None):
    """Resolve ``data_name`` relative to the calling module's package dir."""
    if package is None:
        frame = inspect.stack()[1]
        mod = inspect.getmodule(frame[0])
        base = os.path.dirname(mod.__file__)
    else:
        base = os.path.dirname(sys.modules[package].__file__)
    return os.path.join(base, data_name)
