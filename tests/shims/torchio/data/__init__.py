"""Subject / Image containers of the shim (torchio/data/subject.py, torchio/data/image.py)."""
import copy

import numpy as np
import torch

from ..constants import AFFINE, DATA, INTENSITY, LABEL, PATH, STEM, TYPE

PROTECTED = (DATA, AFFINE, TYPE, PATH, STEM)


class Image(dict):
    def __init__(self, path=None, type=None, tensor=None, affine=None, check_nans=False, **kwargs):  # noqa: A002
        if path is not None:
            raise NotImplementedError("torchio shim: images are built from tensors (no file I/O)")
        if tensor is None:
            raise ValueError("A value for path or tensor must be given")
        if isinstance(tensor, np.ndarray):
            tensor = torch.as_tensor(tensor)
        if not isinstance(tensor, torch.Tensor):
            raise TypeError(f"Input tensor must be a PyTorch tensor or NumPy array, but type {tensor.__class__} was found")
        if tensor.ndim != 4:
            raise ValueError(f"Input tensor must be 4D, but it is {tensor.ndim}D")
        if tensor.dtype == torch.bool:
            tensor = tensor.to(torch.uint8)
        for key in PROTECTED:
            if key in kwargs:
                raise ValueError(f'Key "{key}" is reserved. Use a different one')
        super().__init__(**kwargs)
        self[DATA] = tensor
        self[AFFINE] = np.eye(4) if affine is None else np.asarray(affine, dtype=np.float64)
        self[TYPE] = INTENSITY if type is None else type
        self[PATH] = ""
        self[STEM] = ""
        self.check_nans = check_nans

    def __repr__(self):
        return f"{self.__class__.__name__}(shape: {tuple(self.shape)}; dtype: {self.data.dtype})"

    def __copy__(self):
        extra = {k: copy.deepcopy(v) for k, v in self.items() if k not in PROTECTED}
        return self.__class__(tensor=self.data, affine=self.affine.copy(), type=self[TYPE], **extra)

    @property
    def data(self):
        return self[DATA]

    @data.setter
    def data(self, tensor):
        self.set_data(tensor)

    def set_data(self, tensor):
        if tensor.ndim != 4:
            raise ValueError(f"Input tensor must be 4D, but it is {tensor.ndim}D")
        self[DATA] = tensor

    @property
    def tensor(self):
        return self.data

    @property
    def affine(self):
        return self[AFFINE]

    @affine.setter
    def affine(self, matrix):
        self[AFFINE] = np.asarray(matrix, dtype=np.float64)

    @property
    def type(self):  # noqa: A003
        return self[TYPE]

    @property
    def shape(self):
        return tuple(self.data.shape)

    @property
    def spatial_shape(self):
        return tuple(self.data.shape[1:])

    @property
    def num_channels(self):
        return self.data.shape[0]

    @property
    def spacing(self):
        return tuple(float(np.linalg.norm(self.affine[:3, i])) for i in range(3))

    def numpy(self):
        return np.asarray(self.data)

    def load(self):
        return None


class ScalarImage(Image):
    def __init__(self, *args, **kwargs):
        if "type" in kwargs and kwargs["type"] != INTENSITY:
            raise ValueError("Type of ScalarImage is always torchio.INTENSITY")
        kwargs.update({"type": INTENSITY})
        super().__init__(*args, **kwargs)


class LabelMap(Image):
    def __init__(self, *args, **kwargs):
        if "type" in kwargs and kwargs["type"] != LABEL:
            raise ValueError("Type of LabelMap is always torchio.LABEL")
        kwargs.update({"type": LABEL})
        super().__init__(*args, **kwargs)


def _all_subclasses(cls):
    out = set()
    for sub in cls.__subclasses__():
        out.add(sub)
        out |= _all_subclasses(sub)
    return out


class Subject(dict):
    def __init__(self, *args, **kwargs):
        if args:
            if len(args) == 1 and isinstance(args[0], dict):
                kwargs.update(args[0])
            else:
                raise ValueError("Only one dictionary as positional argument is allowed")
        super().__init__(**kwargs)
        self.applied_transforms = []

    def __repr__(self):
        return f"Subject(Keys: {tuple(self.keys())}; images: {len(self.get_images(intensity_only=False))})"

    def __copy__(self):
        result = {}
        for key, value in self.items():
            result[key] = copy.copy(value) if isinstance(value, Image) else copy.deepcopy(value)
        new = Subject(result)
        new.applied_transforms = self.applied_transforms[:]
        return new

    def __len__(self):
        return len(self.get_images(intensity_only=False))

    def get_images_dict(self, intensity_only=True, include=None, exclude=None):
        images = {}
        for name, image in self.items():
            if not isinstance(image, Image):
                continue
            if intensity_only and not image[TYPE] == INTENSITY:
                continue
            if include is not None and name not in include:
                continue
            if exclude is not None and name in exclude:
                continue
            images[name] = image
        return images

    def get_images(self, intensity_only=True, include=None, exclude=None):
        return list(self.get_images_dict(intensity_only=intensity_only, include=include, exclude=exclude).values())

    def get_first_image(self):
        return self.get_images(intensity_only=False)[0]

    def add_image(self, image, image_name):
        if not isinstance(image, Image):
            raise ValueError(f"Image must be an instance of torchio.Image, not {type(image)}")
        self[image_name] = image

    def remove_image(self, image_name):
        del self[image_name]

    @property
    def shape(self):
        shapes = {im.shape for im in self.get_images(intensity_only=False)}
        if len(shapes) > 1:
            raise RuntimeError(f"More than one shape found in subject images: {shapes}")
        return self.get_first_image().shape

    @property
    def spatial_shape(self):
        shapes = {im.spatial_shape for im in self.get_images(intensity_only=False)}
        if len(shapes) > 1:
            raise RuntimeError(f"More than one spatial shape found in subject images: {shapes}")
        return self.get_first_image().spatial_shape

    def check_consistent_spatial_shape(self):
        return self.spatial_shape

    def load(self):
        return None

    # ---- transform history (torchio/data/subject.py: add_transform / get_applied_transforms / get_composed_history)
    def add_transform(self, transform, parameters_dict):
        self.applied_transforms.append((transform.name, parameters_dict))

    @property
    def history(self):
        return self.get_applied_transforms()

    def get_applied_transforms(self, ignore_intensity=False, image_interpolation=None):
        from ..transforms import IntensityTransform, Transform
        name_to_transform = {cls.__name__: cls for cls in _all_subclasses(Transform)}
        out = []
        for name, arguments in self.applied_transforms:
            transform = name_to_transform[name](**arguments)
            if ignore_intensity and isinstance(transform, IntensityTransform):
                continue
            out.append(transform)
        return out

    def get_composed_history(self, ignore_intensity=False, image_interpolation=None):
        from ..transforms import Compose
        return Compose(self.get_applied_transforms(ignore_intensity=ignore_intensity))

    def get_inverse_transform(self, warn=True, ignore_intensity=True, image_interpolation=None):
        return self.get_composed_history(ignore_intensity=ignore_intensity).inverse(warn=warn)

    def apply_inverse_transform(self, **kwargs):
        transformed = self.get_inverse_transform(**kwargs)(self)
        transformed.clear_history()
        return transformed

    def clear_history(self):
        self.applied_transforms = []


class SubjectsDataset(torch.utils.data.Dataset):
    def __init__(self, subjects, transform=None, load_getitem=True):
        self._subjects = list(subjects)
        self._transform = transform
        self.load_getitem = load_getitem

    def __len__(self):
        return len(self._subjects)

    def __getitem__(self, index):
        subject = copy.deepcopy(self._subjects[int(index)])
        if self._transform is not None:
            subject = self._transform(subject)
        return subject

    def set_transform(self, transform):
        self._transform = transform


def _not_in_shim(name):
    class _Missing:
        def __init__(self, *args, **kwargs):
            raise NotImplementedError(f"torchio.{name} is not part of the test shim: the b200 PatchPredict does its own "
                                      f"patch extraction / aggregation on the device")
    _Missing.__name__ = name
    return _Missing


GridSampler = _not_in_shim("GridSampler")
GridAggregator = _not_in_shim("GridAggregator")
Queue = _not_in_shim("Queue")
