from .random_transform import RandomTransform  # noqa: F401
