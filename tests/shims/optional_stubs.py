"""TEST INFRASTRUCTURE -- import-time stand-ins for third-party packages the reference imports at module level but
that have nothing to do with the hot path and are not installed in this image (matplotlib for contour figures,
skimage for CPU morphology, random_words for run names, ...).  ``install()`` registers a meta-path finder that serves
an inert stub for each missing top-level package (and its submodules); any attempt to USE a stubbed member raises."""
import importlib.abc
import importlib.machinery
import importlib.util
import sys
import types

OPTIONAL = ("matplotlib", "skimage", "random_words", "ipywidgets", "nibabel", "SimpleITK", "fire", "IPython",
            "mpl_toolkits", "PIL", "wandb", "dill", "sklearn", "scipy", "pandas")


class _Inert:
    def __init__(self, name):
        self._name = name

    def __call__(self, *a, **k):
        raise NotImplementedError(f"{self._name} is a test stub (package not installed in this image)")

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return _Inert(f"{self._name}.{item}")

    def __mro_entries__(self, bases):      # usable as a base class in `class X(stub.Base)`
        return (object,)


class _StubModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return _Inert(f"{self.__name__}.{item}")


class _Loader(importlib.abc.Loader):
    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        return None


class StubFinder(importlib.abc.MetaPathFinder):
    def __init__(self, roots):
        self.roots = set(roots)

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in self.roots:
            return importlib.machinery.ModuleSpec(fullname, _Loader(), is_package=True)
        return None


def install(names=OPTIONAL):
    """Stubs every package of ``names`` that cannot be found; returns the list of stubbed names."""
    missing = []
    for name in names:
        if name in sys.modules:
            continue
        try:
            found = importlib.util.find_spec(name) is not None
        except (ImportError, ValueError):
            found = False
        if not found:
            missing.append(name)
    if missing:
        sys.meta_path.append(StubFinder(missing))
    return missing
