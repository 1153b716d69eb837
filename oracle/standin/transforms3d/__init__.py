"""Stand-in for the few transforms3d entry points the reference calls.
Written from the published algorithms (Gohlke / transforms3d docs); test
infrastructure only, see ../README.md."""
from . import affines, axangles, utils, euler, quaternions  # noqa: F401
