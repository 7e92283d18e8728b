from .label_transform import LabelTransform  # noqa: F401
