from . import label, spatial  # noqa: F401
