from abc import ABC, abstractmethod

from ..utils import auto_str


class Evaluator(ABC):
    """``__call__(subjects) -> dict`` (reference evaluators/evaluator.py:9-15)."""

    @abstractmethod
    def __call__(self, subjects) -> dict:
        raise NotImplementedError()

    def __repr__(self):
        return auto_str(self)
