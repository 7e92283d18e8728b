import warnings

import numpy as np
import torch

from .spatial_transform import SpatialTransform
from .transform import Transform


class Compose(Transform):
    def __init__(self, transforms, **kwargs):
        super().__init__(**kwargs)
        for transform in transforms:
            if not callable(transform):
                raise TypeError(f"The object {transform} of type {type(transform)} is not callable")
        self.transforms = list(transforms)

    def __len__(self):
        return len(self.transforms)

    def __getitem__(self, index):
        return self.transforms[index]

    def __repr__(self):
        return f"{self.name}({self.transforms})"

    def apply_transform(self, subject):
        for transform in self.transforms:
            subject = transform(subject)
        return subject

    def is_invertible(self):
        return all(t.is_invertible() for t in self.transforms)

    def inverse(self, warn=True):
        transforms = []
        for transform in self.transforms:
            if transform.is_invertible():
                transforms.append(transform.inverse())
            elif warn:
                warnings.warn(f"Skipping {transform.name} as it is not invertible", RuntimeWarning)
        transforms.reverse()
        result = Compose(transforms)
        if not transforms and warn:
            warnings.warn("No invertible transforms found", RuntimeWarning)
        return result


class _Bounds(SpatialTransform):
    def __init__(self, bounds_parameters, **kwargs):
        super().__init__(**kwargs)
        self.bounds_parameters = self.parse_bounds(bounds_parameters)

    @staticmethod
    def parse_bounds(b):
        if isinstance(b, int):
            return (b,) * 6
        b = tuple(int(v) for v in b)
        if len(b) == 3:
            return (b[0], b[0], b[1], b[1], b[2], b[2])
        if len(b) != 6:
            raise ValueError(f"Bounds must be an int or a tuple of 3 or 6 ints, not {b}")
        return b


class Pad(_Bounds):
    def __init__(self, padding, padding_mode=0, **kwargs):
        super().__init__(padding, **kwargs)
        self.padding = padding
        self.padding_mode = padding_mode
        self.args_names = ("padding", "padding_mode")

    def apply_transform(self, subject):
        lo = np.array(self.bounds_parameters[::2])
        pads = ((0, 0),) + tuple(zip(self.bounds_parameters[::2], self.bounds_parameters[1::2]))
        for image in self.get_images(subject):
            affine = image.affine.copy()
            affine[:3, 3] = (affine @ np.array([*(-lo), 1.0]))[:3]
            if isinstance(self.padding_mode, (int, float)):
                padded = np.pad(image.numpy(), pads, mode="constant", constant_values=self.padding_mode)
            else:
                padded = np.pad(image.numpy(), pads, mode=self.padding_mode)
            image.set_data(torch.as_tensor(padded))
            image.affine = affine
        return subject

    def is_invertible(self):
        return True

    def inverse(self):
        return Crop(self.padding)


class Crop(_Bounds):
    def __init__(self, cropping, **kwargs):
        super().__init__(cropping, **kwargs)
        self.cropping = cropping
        self.args_names = ("cropping",)

    def apply_transform(self, subject):
        i0, i1, j0, j1, k0, k1 = self.bounds_parameters
        for image in self.get_images(subject):
            w, h, d = image.spatial_shape
            affine = image.affine.copy()
            affine[:3, 3] = (affine @ np.array([i0, j0, k0, 1.0]))[:3]
            image.set_data(image.data[:, i0:w - i1, j0:h - j1, k0:d - k1].clone())
            image.affine = affine
        return subject

    def is_invertible(self):
        return True

    def inverse(self):
        return Pad(self.cropping)


class CopyAffine(SpatialTransform):
    def __init__(self, target, **kwargs):
        super().__init__(**kwargs)
        if not isinstance(target, str):
            raise ValueError(f"The target must be a string, but \"{type(target)}\" was found")
        self.target = target
        self.args_names = ("target",)

    def apply_transform(self, subject):
        if self.target not in subject:
            raise RuntimeError(f"Target image \"{self.target}\" not found in subject")
        affine = subject[self.target].affine
        for image in self.get_images(subject):
            image.affine = affine.copy()
        return subject
