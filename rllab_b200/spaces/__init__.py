from .box import Box  # noqa: F401
