"""Import pieces of the UNMODIFIED reference from /root/reference (oracle; test infrastructure only).

Only usable in the build container: /root/reference does not exist on the GPU box, so nothing in
``-m gpu`` tests, ``smoke()`` or ``bench.py`` calls into this module.  It exists to (a) generate the
committed golden vectors (``oracle/make_golden.py``) and (b) cross-check the restatements in the
``-m "not gpu"`` tests when the reference happens to be present.

Nothing is copied: the reference modules are imported from where they lie, with stub modules standing
in for packages that are absent here (matplotlib; torch_mimicry).
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = os.environ.get("SDG_REFERENCE_ROOT", "/root/reference")
REF_PKG = os.path.join(REF_ROOT, "diagan-pkg")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_PKG, "diagan"))


def _stub(name: str, **attrs):
    if name in sys.modules:
        mod = sys.modules[name]
    else:
        mod = types.ModuleType(name)
        sys.modules[name] = mod
    for k, v in attrs.items():
        setattr(mod, k, v)
    return mod


def _ensure_path():
    if not available():
        raise RuntimeError(f"reference not present at {REF_ROOT}")
    if REF_PKG not in sys.path:
        sys.path.insert(0, REF_PKG)


def _stub_matplotlib():
    try:
        import matplotlib  # noqa: F401
        return
    except Exception:
        pass
    mpl = _stub("matplotlib")
    plt = _stub("matplotlib.pyplot")
    mpl.pyplot = plt
    mpl.use = lambda *a, **k: None


def _stub_mimicry():
    try:
        import torch_mimicry  # noqa: F401
        return
    except Exception:
        pass
    import torch.nn as nn

    class BaseDiscriminator(nn.Module):
        def __init__(self, ndf=None, loss_type=None, **kw):
            super().__init__()
            self.ndf, self.loss_type = ndf, loss_type

    class BaseGenerator(nn.Module):
        def __init__(self, nz=None, ngf=None, bottom_width=None, loss_type=None, **kw):
            super().__init__()
            self.nz, self.ngf, self.bottom_width, self.loss_type = nz, ngf, bottom_width, loss_type

    def _loss(*a, **k):
        raise NotImplementedError("torch_mimicry stub")

    class _Placeholder(nn.Module):
        pass

    root = _stub("torch_mimicry")
    nets = _stub("torch_mimicry.nets")
    gan = _stub("torch_mimicry.nets.gan")
    gan_gan = _stub("torch_mimicry.nets.gan.gan", BaseDiscriminator=BaseDiscriminator, BaseGenerator=BaseGenerator)
    modules = _stub("torch_mimicry.modules")
    losses = _stub("torch_mimicry.modules.losses", hinge_loss_dis=_loss, minimax_loss_dis=_loss,
                   hinge_loss_gen=_loss, minimax_loss_gen=_loss, ns_loss_gen=_loss)
    sn = {n: type(n, (_Placeholder,), {}) for n in
          ("SNGANGenerator32", "SNGANGenerator64", "SNGANDiscriminator32", "SNGANDiscriminator64")}
    info = {n: type(n, (_Placeholder,), {}) for n in
            ("InfoMaxGANGenerator32", "InfoMaxGANGenerator64", "InfoMaxGANDiscriminator32", "InfoMaxGANDiscriminator64")}
    ss = {n: type(n, (_Placeholder,), {}) for n in
          ("SSGANGenerator32", "SSGANGenerator64", "SSGANDiscriminator32", "SSGANDiscriminator64")}
    sngan = _stub("torch_mimicry.nets.sngan", **sn)
    infomax = _stub("torch_mimicry.nets.infomax_gan", **info)
    ssgan = _stub("torch_mimicry.nets.ssgan", **ss)
    root.nets, root.modules = nets, modules
    nets.gan, nets.sngan, nets.infomax_gan, nets.ssgan = gan, sngan, infomax, ssgan
    gan.gan = gan_gan
    modules.losses = losses


def reference_calculate_scores():
    """-> the reference's own ``diagan.utils.plot.calculate_scores`` (plot.py:220-249)."""
    _ensure_path()
    _stub_matplotlib()
    from diagan.utils.plot import calculate_scores
    return calculate_scores


def reference_drs_class():
    """-> the reference's own ``diagan.models.drs.DRS`` (drs.py:10-69)."""
    _ensure_path()
    _stub_mimicry()
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_drs", os.path.join(REF_PKG, "diagan", "models", "drs.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.DRS


def reference_get_transform():
    """``diagan.datasets.transform.get_transform`` (transform.py:35-41); needs torchvision + Pillow (installed here)."""
    _ensure_path()
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_transform", os.path.join(REF_PKG, "diagan", "datasets", "transform.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.get_transform


def reference_dcgan_discriminator():
    """-> the reference's own ``MNIST_DCGAN_Discriminator`` class (mnist.py:155-223)."""
    _ensure_path()
    _stub_matplotlib()
    _stub_mimicry()
    from diagan.models.mnist import MNIST_DCGAN_Discriminator
    return MNIST_DCGAN_Discriminator


def reference_stylegan2_discriminator():
    """-> the reference's own ``StyleGANDiscriminator`` class (diagan/models/stylegan2.py:619-677).

    Importing ``diagan.models.op`` JIT-compiles two CUDA extensions at import time (op/fused_act.py:11-17,
    op/upfirdn2d.py:10-16).  The oracle only needs the CPU fall-backs of those ops (fused_act.py:104-116,
    upfirdn2d.py:145-200), so the JIT loader is stubbed out for the duration of the import."""
    _ensure_path()
    _stub_matplotlib()
    _stub_mimicry()
    import torch.utils.cpp_extension as ext
    real = ext.load
    ext.load = lambda *a, **k: None
    try:
        from diagan.models.stylegan2 import StyleGANDiscriminator
    finally:
        ext.load = real
    return StyleGANDiscriminator


class _PermissiveModule(types.ModuleType):
    """Stand-in for a package that is absent here: any attribute is a fresh empty class (usable as a base class, never
    called on the diagnosis path), any submodule imports.  Explicit stubs set earlier win."""
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        full = f"{self.__name__}.{name}"
        if full in sys.modules:
            return sys.modules[full]
        obj = type(name, (), {})
        setattr(self, name, obj)
        return obj


class _PermissiveFinder:
    def __init__(self, roots):
        self.roots = tuple(roots)

    def find_spec(self, fullname, path=None, target=None):
        import importlib.machinery
        if fullname.split(".")[0] in self.roots:
            return importlib.machinery.ModuleSpec(fullname, self)
        return None

    def create_module(self, spec):
        return _PermissiveModule(spec.name)

    def exec_module(self, module):
        pass


def stub_training_stack():
    """Make ``diagan.trainer.trainer`` / ``diagan.trainer.evaluate`` importable here: torch_mimicry (incl. ``.training``,
    ``.metrics``, ``.utils``) and tensorflow are absent, neither is touched by the diagnosis path itself.  Only for the
    drop-in tests (tests/test_dropin_cpu.py): proves ``diagan_b200.patch.install()`` binds onto the REAL classes."""
    _ensure_path()
    _stub_matplotlib()
    _stub_mimicry()
    import numpy as np
    if "numpy.lib.type_check" not in sys.modules:          # removed in NumPy 2 (compute_fid_with_attr.py:9 imports `imag`)
        try:
            import numpy.lib.type_check  # noqa: F401
        except Exception:
            _stub("numpy.lib.type_check", imag=np.imag, real=np.real)
    missing = []
    for root in ("tensorflow", "torch_mimicry", "lmdb", "imageio", "cv2", "tensorboard", "tensorboardX", "seaborn"):
        mod = sys.modules.get(root)
        if mod is None or not getattr(mod, "__file__", None):
            missing.append(root)
    if missing and not any(isinstance(f, _PermissiveFinder) for f in sys.meta_path):
        sys.meta_path.append(_PermissiveFinder(missing))
    if "torch_mimicry" in missing:
        # packages stubbed as plain modules by _stub_mimicry need a __path__ so that their submodules import
        for name, mod in list(sys.modules.items()):
            if name.split(".")[0] == "torch_mimicry" and not hasattr(mod, "__path__"):
                mod.__path__ = []
        import importlib
        tr = importlib.import_module("torch_mimicry.training")
        tr.Trainer = type("Trainer", (), {})
        for sub in ("logger", "metric_log"):
            setattr(tr, sub, importlib.import_module(f"torch_mimicry.training.{sub}"))
        sys.modules["torch_mimicry"].training = tr
