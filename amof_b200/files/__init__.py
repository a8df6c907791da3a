from . import path  # noqa: F401
