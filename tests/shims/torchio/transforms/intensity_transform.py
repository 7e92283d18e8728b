from ..data import ScalarImage
from .transform import Transform


class IntensityTransform(Transform):
    """Transform that modifies voxel intensities only (applied to ScalarImage instances)."""

    def get_images(self, subject):
        return [im for im in super().get_images(subject) if isinstance(im, ScalarImage)]

    def get_images_dict(self, subject):
        return {k: v for k, v in super().get_images_dict(subject).items() if isinstance(v, ScalarImage)}
