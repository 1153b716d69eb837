from . import ascii  # noqa: F401
