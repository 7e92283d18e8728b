from ..transform import Transform


class RandomTransform(Transform):
    """Base class for stochastic augmentation transforms."""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)

    def add_include_exclude(self, kwargs):
        kwargs["include"] = self.include
        kwargs["exclude"] = self.exclude
        return kwargs
