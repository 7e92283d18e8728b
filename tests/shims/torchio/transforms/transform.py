"""torchio/transforms/transform.py (0.18.x), reduced to Subject inputs."""
import copy
import numbers
from abc import ABC, abstractmethod
from typing import Callable, Optional, Sequence, Union

import torch

from ..data import Image, Subject

TypeMaskingMethod = Union[str, Callable, int, Sequence[int], None]


class Transform(ABC):
    def __init__(self, p: float = 1, copy: bool = True, include=None, exclude=None, keys=None, keep=None):  # noqa: A002
        self.probability = self.parse_probability(p)
        self.copy = copy
        if keys is not None:
            include = keys
        self.include, self.exclude = self.parse_include_and_exclude(include, exclude)
        self.keep = keep
        self.args_names = ()

    def __call__(self, data):
        if torch.rand(1).item() > self.probability:
            return data
        if not isinstance(data, Subject):
            raise NotImplementedError("torchio shim: transforms take Subject instances")
        subject = data
        if self.keep is not None:
            images_to_keep = {new: copy.copy(subject[name]) for name, new in self.keep.items()}
        if self.copy:
            subject = copy.copy(subject)
        transformed = self.apply_transform(subject)
        if self.keep is not None:
            for name, image in images_to_keep.items():
                transformed.add_image(image, name)
        self.add_transform_to_subject_history(transformed)
        return transformed

    def __repr__(self):
        if hasattr(self, "args_names"):
            names = self.args_names
            args_strings = [f"{arg}={getattr(self, arg)}" for arg in names]
            return f"{self.name}({', '.join(args_strings)})"
        return super().__repr__()

    @property
    def name(self):
        return self.__class__.__name__

    @abstractmethod
    def apply_transform(self, subject):
        raise NotImplementedError

    def add_transform_to_subject_history(self, subject):
        from .augmentation import RandomTransform
        from .compose import Compose
        if not isinstance(self, (RandomTransform, Compose)):
            subject.add_transform(self, self._get_reproducing_arguments())

    def _get_reproducing_arguments(self):
        reproducing_arguments = {"include": self.include, "exclude": self.exclude, "copy": self.copy}
        reproducing_arguments.update({name: getattr(self, name) for name in self.args_names})
        return reproducing_arguments

    def is_invertible(self):
        return hasattr(self, "invert_transform")

    def inverse(self):
        if not self.is_invertible():
            raise RuntimeError(f"{self.name} is not invertible")
        new = copy.deepcopy(self)
        new.invert_transform = not self.invert_transform
        return new

    @staticmethod
    def parse_probability(probability):
        is_number = isinstance(probability, numbers.Number)
        if not (is_number and 0 <= probability <= 1):
            raise ValueError(f"Probability must be a number in [0, 1], not {probability}")
        return probability

    @staticmethod
    def validate_keys_sequence(keys, name):
        if keys is None:
            return
        if isinstance(keys, str):
            raise ValueError(f'"{name}" must be a sequence of strings, not a string "{keys}"')
        if not isinstance(keys, (list, tuple)):
            raise ValueError(f'"{name}" must be a sequence of strings, not {type(keys)}')

    def parse_include_and_exclude(self, include=None, exclude=None):
        if include is not None and exclude is not None:
            raise ValueError("Include and exclude cannot both be specified")
        self.validate_keys_sequence(include, "include")
        self.validate_keys_sequence(exclude, "exclude")
        return include, exclude

    def get_images(self, subject):
        return subject.get_images(intensity_only=False, include=self.include, exclude=self.exclude)

    def get_images_dict(self, subject):
        return subject.get_images_dict(intensity_only=False, include=self.include, exclude=self.exclude)

    @staticmethod
    def get_mask_from_masking_method(masking_method, subject, tensor, labels=None):
        if masking_method is None:
            return torch.ones_like(tensor, dtype=torch.bool)
        if callable(masking_method):
            return masking_method(tensor)
        if isinstance(masking_method, str) and masking_method in subject:
            data = subject[masking_method].data
            if labels is None:
                return data.bool()
            mask = torch.zeros_like(data, dtype=torch.bool)
            for label in labels:
                mask |= data == label
            return mask
        raise NotImplementedError(f"torchio shim: masking method {masking_method!r}")
