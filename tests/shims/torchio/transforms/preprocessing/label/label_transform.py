from ....data import LabelMap
from ...transform import Transform


class LabelTransform(Transform):
    """Transform that modifies label maps only."""

    def get_images(self, subject):
        return [im for im in super().get_images(subject) if isinstance(im, LabelMap)]

    def get_images_dict(self, subject):
        return {k: v for k, v in super().get_images_dict(subject).items() if isinstance(v, LabelMap)}
