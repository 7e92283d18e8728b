from .transform import Transform, TypeMaskingMethod  # noqa: F401
from .spatial_transform import SpatialTransform  # noqa: F401
from .intensity_transform import IntensityTransform  # noqa: F401
from .augmentation import RandomTransform  # noqa: F401
from .preprocessing.label.label_transform import LabelTransform  # noqa: F401
from .preprocessing.spatial.resample import Resample  # noqa: F401
from .compose import Compose, CopyAffine, Crop, Pad  # noqa: F401
from . import preprocessing  # noqa: F401
