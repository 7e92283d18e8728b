from typing import Tuple, Union

from ...spatial_transform import SpatialTransform

TypeSpacing = Union[float, Tuple[float, float, float]]


class Resample(SpatialTransform):
    def __init__(self, target=1, image_interpolation="linear", pre_affine_name=None, scalars_only=False, **kwargs):
        super().__init__(**kwargs)
        self.target = target
        self.image_interpolation = image_interpolation
        self.pre_affine_name = pre_affine_name
        self.scalars_only = scalars_only
        self.args_names = ("target", "image_interpolation", "pre_affine_name", "scalars_only")

    @staticmethod
    def parse_spacing(spacing):
        if isinstance(spacing, (tuple, list)) and len(spacing) == 3:
            result = tuple(float(s) for s in spacing)
        elif isinstance(spacing, (int, float)):
            result = 3 * (float(spacing),)
        else:
            raise ValueError(f"Target must be a string, a positive number or a sequence of positive numbers, not {type(spacing)}")
        if any(s <= 0 for s in result):
            raise ValueError(f"Spacing must be strictly positive, not \"{spacing}\"")
        return result

    def apply_transform(self, subject):
        raise NotImplementedError("torchio shim: Resample needs SimpleITK and is outside the tested path")
