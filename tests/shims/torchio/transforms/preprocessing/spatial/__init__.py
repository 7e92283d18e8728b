from .resample import Resample  # noqa: F401
