"""Import files of an UNMODIFIED hailanyi/CPD checkout on top of cpd_b200 ("import closure").

`cpd/__init__`, `cpd/datasets/__init__`, `cpd/models/__init__` ... pull the whole project in (datasets, CLI config,
four more compiled extensions, easydict, skimage, prefetch_generator, tensorboardX), none of which the hot path
needs.  ``install_reference(root)``

  * calls ``compat.install()`` (spconv / cumm namespaces backed by libcpd_b200.so),
  * registers every directory under ``root/cpd`` as a BARE package (its ``__path__`` only -- the ``__init__.py`` files
    are not executed), so ``import cpd.models.backbones_3d.spconv_backbone`` executes exactly that file and the files
    it imports itself,
  * maps ``cpd.ops.iou3d_nms.iou3d_nms_cuda`` to cpd_b200.iou3d_nms_cuda,
  * registers import stubs for the compiled extensions that are out of scope (they raise if CALLED) and tiny
    stand-ins for the third-party packages the reference imports at module level but never uses on the hot path.

Nothing of the reference is copied or patched; ``uninstall_reference()`` removes every entry again.
"""
import importlib
import os
import sys
import types

from . import install as _install_shim

_REGISTERED = []

# compiled extensions of the reference that are NOT part of the hot path
_OUT_OF_SCOPE_EXTENSIONS = (
    "cpd.ops.roiaware_pool3d.roiaware_pool3d_cuda",
    "cpd.ops.roipoint_pool3d.roipoint_pool3d_cuda",
    "cpd.ops.pointnet2.pointnet2_batch.pointnet2_batch_cuda",
    "cpd.ops.votr_ops.votr_ops_cuda",
    "cpd.ops.dcn.deform_conv_cuda",
)


class _Unavailable(types.ModuleType):
    """Import stub: importing works, using anything raises."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        raise ImportError(f"{self.__name__}.{name}: this extension / package is outside the cpd_b200 hot path (import stub)")


class EasyDict(dict):
    """Stand-in for easydict.EasyDict (attribute access, recursive)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            v = EasyDict(v)
        elif isinstance(v, (list, tuple)):
            v = type(v)(EasyDict(x) if isinstance(x, dict) and not isinstance(x, EasyDict) else x for x in v)
        super().__setitem__(k, v)

    __setattr__ = __setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


class BackgroundGenerator:
    """Stand-in for prefetch_generator.BackgroundGenerator: iterates the wrapped iterable in the calling thread."""

    def __init__(self, generator, max_prefetch=1):
        self.generator = generator

    def __iter__(self):
        return iter(self.generator)


def _register(name, module):
    if name not in sys.modules:
        sys.modules[name] = module
        _REGISTERED.append(name)
    return sys.modules[name]


def install_reference(root):
    """root: path of the hailanyi/CPD checkout (the directory that contains ``cpd/``)."""
    pkg_root = os.path.join(root, "cpd")
    if not os.path.isdir(pkg_root):
        raise FileNotFoundError(f"{pkg_root} is not a CPD checkout")
    _install_shim()
    for dirpath, dirnames, filenames in os.walk(pkg_root):
        dirnames[:] = [d for d in dirnames if not d.startswith((".", "__")) and d != "src"]
        rel = os.path.relpath(dirpath, root).replace(os.sep, ".")
        mod = types.ModuleType(rel)
        mod.__path__ = [dirpath]
        mod.__package__ = rel
        mod.__cpd_b200_bare__ = True
        _register(rel, mod)
        parent, _, leaf = rel.rpartition(".")
        if parent and parent in sys.modules:
            setattr(sys.modules[parent], leaf, sys.modules[rel])
    from .. import iou3d_nms_cuda
    _register("cpd.ops.iou3d_nms.iou3d_nms_cuda", iou3d_nms_cuda)
    sys.modules["cpd.ops.iou3d_nms"].iou3d_nms_cuda = iou3d_nms_cuda
    from .. import pointnet2_stack_cuda           # voxel query + grouping of the RoI grid pooling
    _register("cpd.ops.pointnet2.pointnet2_stack.pointnet2_stack_cuda", pointnet2_stack_cuda)
    sys.modules["cpd.ops.pointnet2.pointnet2_stack"].pointnet2_stack_cuda = pointnet2_stack_cuda
    for name in _OUT_OF_SCOPE_EXTENSIONS:
        stub = _register(name, _Unavailable(name))
        parent, _, leaf = name.rpartition(".")
        if parent in sys.modules:
            setattr(sys.modules[parent], leaf, stub)
    # third-party packages the reference imports at module level; real ones win when installed
    for name, attrs in (("easydict", dict(EasyDict=EasyDict)), ("prefetch_generator", dict(BackgroundGenerator=BackgroundGenerator))):
        try:
            importlib.import_module(name)
        except ImportError:
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            _register(name, m)
    for name in ("skimage", "skimage.transform", "skimage.io", "tensorboardX", "SharedArray"):
        try:
            importlib.import_module(name)
        except ImportError:
            stub = _register(name, _Unavailable(name))
            parent, _, leaf = name.rpartition(".")
            if parent in sys.modules:
                setattr(sys.modules[parent], leaf, stub)
    return sys.modules["cpd"]


def uninstall_reference():
    for name in list(sys.modules):
        if name in _REGISTERED or ((name == "cpd" or name.startswith("cpd.")) and name in sys.modules):
            m = sys.modules[name]
            if name in _REGISTERED or getattr(m, "__file__", "") or getattr(m, "__cpd_b200_bare__", False):
                del sys.modules[name]
    _REGISTERED.clear()
