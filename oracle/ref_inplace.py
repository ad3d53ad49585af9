"""TEST INFRASTRUCTURE ONLY -- in-place import of the *unmodified* reference (read-only /root/reference).

Used in the build container to (a) validate ``oracle/refmath.py`` against the real reference code and
(b) generate the golden vectors under ``tests/golden/`` (``oracle/gen_golden.py``).  It cannot travel to
the GPU box (``/root/reference`` does not exist there), so nothing under ``-m gpu``, ``smoke()`` or
``bench.py`` imports this module.

Shims applied -- none of them changes the arithmetic of the hot path:
  * empty ``sys.modules`` stubs for optional packages the reference imports at module load time
    (h5py, imageio, seaborn, matplotlib, wget, torchnet, pytorch_fid, cv2 ...);
  * ``np.product = np.prod`` (removed in numpy 2; used by reference decoders/encoders);
  * CPU runs: ``Tensor.cuda`` / ``Module.cuda`` identity and ``.to("cuda")`` -> no-op, because the reference
    hard-codes CUDA placement (mmvae_models.py:45,51,173,222,249,309,... objectives.py:165,...,500);
  * N1: ``MultimodalObjective.iwae`` calls ``data["pz_params"].cuda()`` on a tuple
    (objectives.py:353) -> we pass a tuple subclass with a ``.cuda()`` method, the arithmetic is untouched;
  * eps injection: ``torch.distributions.normal._standard_normal`` / ``Laplace.rsample`` are replaced by
    functions popping caller supplied noise (same formulas as torch: normal.py:82-101, laplace.py:73-90).
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

REF_ROOT = os.environ.get("MMVAE_REFERENCE_ROOT", "/root/reference/multimodal_compare")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "models"))


class _Anything:
    """Attribute sink for stubbed optional modules."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, item):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    try:
        return importlib.import_module(name)
    except Exception:
        pass
    m = types.ModuleType(name)
    m.__dict__.update(attrs)

    def _ga(item):
        if item.startswith("__"):
            raise AttributeError(item)
        return _Anything

    m.__getattr__ = _ga  # type: ignore[attr-defined]
    m.__path__ = []  # behave as a package so that "import a.b" works
    sys.modules[name] = m
    return m


_loaded = None


def load(cpu_shims: bool = True):
    """Import the reference ``models`` package in place and return (models, objectives, utils)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    if not hasattr(np, "product"):
        np.product = np.prod
    import torchvision  # noqa: F401  (real module, must load before any stub is registered)
    for name in ["h5py", "imageio", "seaborn", "matplotlib", "matplotlib.colors", "matplotlib.pyplot",
                 "matplotlib.patches", "matplotlib.cm", "wget", "torchnet", "torchnet.dataset", "pytorch_fid",
                 "pytorch_fid.inception", "cv2", "umap", "pytorch_lightning", "adabelief_pytorch",
                 "gym", "gymnasium", "pybullet", "glob2", "sklearn.manifold", "PIL", "PIL.Image"]:
        _stub(name)
    sys.modules["pytorch_fid.inception"].InceptionV3 = type("InceptionV3", (), {"BLOCK_INDEX_BY_DIM": {2048: 3}})
    if cpu_shims and not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        _orig_to = torch.Tensor.to

        def _to(self, *a, **k):
            a = tuple("cpu" if (isinstance(x, str) and x.startswith("cuda")) else x for x in a)
            if isinstance(k.get("device"), str) and k["device"].startswith("cuda"):
                k["device"] = "cpu"
            return _orig_to(self, *a, **k)

        torch.Tensor.to = _to
        _orig_mto = torch.nn.Module.to

        def _mto(self, *a, **k):
            a = tuple("cpu" if (isinstance(x, str) and x.startswith("cuda")) else x for x in a)
            return _orig_mto(self, *a, **k)

        torch.nn.Module.to = _mto
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    models = importlib.import_module("models")
    objectives = importlib.import_module("models.objectives")
    utils = importlib.import_module("utils")
    _loaded = (models, objectives, utils)
    return _loaded


class PzParams(tuple):
    """N1 shim: a tuple that answers ``.cuda()`` (objectives.py:353 calls it on the pz_params tuple)."""

    def cuda(self):
        return PzParams(x.cuda() for x in self)


class NoiseInjector:
    """Context manager feeding pre-generated noise to every ``rsample`` of Normal / Laplace.

    Normal.rsample = loc + _standard_normal(shape) * scale (torch normal.py:82-86): we replace
    ``_standard_normal``.  Laplace.rsample draws u ~ U(finfo.eps-1, 1) and returns
    loc - scale*sign(u)*log1p(-|u|) (torch laplace.py:73-84): we replace the method with the same
    formula on a supplied ``u``.
    """

    def __init__(self, noises):
        self.noises = list(noises)
        self.log = []

    def _pop(self, shape):
        if not self.noises:
            raise RuntimeError("NoiseInjector ran out of noise tensors (wanted %s)" % (tuple(shape),))
        e = self.noises.pop(0)
        assert tuple(e.shape) == tuple(shape), "noise shape %s != requested %s" % (tuple(e.shape), tuple(shape))
        self.log.append(tuple(shape))
        return e

    def __enter__(self):
        import torch.distributions.normal as tn
        import torch.distributions.laplace as tl
        self._tn, self._tl = tn, tl
        self._orig_sn = tn._standard_normal
        self._orig_lr = tl.Laplace.rsample
        inj = self

        def sn(shape, dtype, device):
            return inj._pop(shape).to(dtype)

        def lr(self_, sample_shape=torch.Size()):
            shape = self_._extended_shape(sample_shape)
            u = inj._pop(shape)
            return self_.loc - self_.scale * u.sign() * torch.log1p(-u.abs())

        tn._standard_normal = sn
        tl.Laplace.rsample = lr
        return self

    def __exit__(self, *exc):
        self._tn._standard_normal = self._orig_sn
        self._tl.Laplace.rsample = self._orig_lr
        return False
