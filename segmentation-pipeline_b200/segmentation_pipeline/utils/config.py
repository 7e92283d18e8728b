"""``Config`` mixin -- mirror of the reference's utils/config.py:9-62: every ``__init__`` argument must be
stored as an attribute of the same name so that ``get_config`` can read it back."""
from __future__ import annotations

from inspect import signature
from numbers import Number
from typing import Any, Dict

from .utils import is_sequence


def get_nested_config(elem):
    if isinstance(elem, Dict):
        return {k: get_nested_config(v) for k, v in elem.items()}
    if is_sequence(elem):
        return [get_nested_config(v) for v in elem]
    if isinstance(elem, Config):
        return get_nested_config(elem.get_config())
    if isinstance(elem, (Number, str, bool)):
        return elem
    return str(elem)


class Config:
    def get_config(self) -> Dict[str, Any]:
        names = list(signature(self.__init__).parameters.keys())
        missing = [n for n in names if n not in self.__dict__]
        if missing:
            raise RuntimeError(f"All parameters for __init__ must be saved as class properties with the same name "
                               f"in order to use default get_config(). The parameter {missing[0]} was not saved.")
        return {n: self.__dict__[n] for n in names}

    def get_nested_config(self) -> Dict[str, Any]:
        return get_nested_config(self)

    def __repr__(self) -> str:
        args = ", ".join(f"{k}={v}" for k, v in self.get_config().items())
        return f"{self.__class__.__name__}({args})"
