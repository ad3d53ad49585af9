"""Import shim: ``import mmvae_b200`` loads the package that lives in ``multimodal-vae-comparison_b200/``
(a directory name Python cannot import directly because of the hyphens)."""
import importlib.util
import os
import sys

_d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multimodal-vae-comparison_b200")
_spec = importlib.util.spec_from_file_location("mmvae_b200", os.path.join(_d, "__init__.py"),
                                               submodule_search_locations=[_d])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["mmvae_b200"] = _mod
_spec.loader.exec_module(_mod)
