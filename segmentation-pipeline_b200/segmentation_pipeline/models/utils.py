"""``filter_kwargs`` -- mirror of the reference's models/utils.py:4-9."""
from inspect import signature


def filter_kwargs(constructor, **kwargs):
    """Keeps only the keyword arguments that ``constructor`` accepts."""
    accepted = signature(constructor).parameters
    return {name: value for name, value in kwargs.items() if name in accepted}
