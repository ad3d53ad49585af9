"""CPU, build container only (skipped where /root/reference is absent): the drop-in plugin classes accept the REFERENCE's
own ``models.vae.VAE`` objects (real FNN encoders / decoders built by the reference's DencoderFactory), exactly as
reference trainer.py:99-111 constructs them -- registry lookup by ``cfg.mixing.lower()``, constructor signature,
``isinstance(model, TorchMMVAE)``, state-dict keys (reference checkpoints load unchanged), ``llik_scaling: auto``,
``pz_params`` and the ``encode()`` dict layout.  No kernel is launched here (no GPU): the arithmetic of the plugins is
covered by tests/test_models_gpu.py against the frozen outputs of the same reference classes."""
import pytest
import torch
import torch.nn as nn

from oracle import ref_inplace

pytestmark = pytest.mark.skipif(not ref_inplace.available(), reason="needs the reference checkout (/root/reference)")

FEATURE_DIMS = {"mod_1": [3, 8, 8], "mod_2": [5, 27]}


def _reference_vaes(models, private=None, llik=("auto", 1)):
    from models.vae import VAE  # the reference's class (oracle/ref_inplace.py put its root on sys.path)
    vaes = {}
    for i, (name, fd) in enumerate(FEATURE_DIMS.items()):
        vaes[name] = VAE("FNN", "FNN", fd, 6, "bce" if i == 0 else "category_ce", private, obj_fn="elbo", beta=1,
                         id_name=name, prior_dist="normal", post_dist="normal", likelihood_dist="normal",
                         llik_scaling=llik[i])
    return vaes


@pytest.mark.parametrize("mixing", ["poe", "moe", "mopoe", "dmvae"])
def test_plugins_wrap_reference_vae_objects(mixing):
    import mmvae_b200
    models, _, _ = ref_inplace.load()
    private = 3 if mixing == "dmvae" else None
    obj_cfg = {"obj": "elbo", "beta": 1.0, "K": 1}
    ref_model = getattr(models, mixing)(nn.ModuleDict(_reference_vaes(models, private)), 6, obj_cfg, None)
    ours = getattr(mmvae_b200, mixing)(nn.ModuleDict(_reference_vaes(models, private)), 6, obj_cfg, None)
    assert isinstance(ours, mmvae_b200.TorchMMVAE) and ours.modelName == ref_model.modelName
    # checkpoints: identical parameter / buffer names and shapes, and a reference state dict loads strictly
    sd_ref, sd_ours = ref_model.state_dict(), ours.state_dict()
    assert list(sd_ref.keys()) == list(sd_ours.keys())
    assert all(sd_ref[k].shape == sd_ours[k].shape for k in sd_ref)
    ours.load_state_dict(sd_ref, strict=True)
    # llik_scaling: auto -> min_dim / dim (mmvae_base.py:41-47)
    for name in FEATURE_DIMS:
        assert float(ours.vaes[name].llik_scaling) == pytest.approx(float(ref_model.vaes[name].llik_scaling))
    assert ours.vaes["mod_1"].llik_scaling == pytest.approx(135.0 / 192.0)
    # prior parameters: (mu0, softmax(logits) * D), same trainable flags
    mu0, s0 = ours.pz_params
    assert mu0.shape == (1, 6) and torch.allclose(s0, torch.ones(1, 6))
    assert [p.requires_grad for p in ours._pz_params] == [p.requires_grad for p in ref_model._pz_params]
    assert ours.latent_factorization == ref_model.latent_factorization
    # encode(): {mod: {"shared": (mu, s), "private": ...}} with the reference encoders' outputs
    g = torch.Generator().manual_seed(0)
    batch = {"mod_1": {"data": torch.rand(4, 3, 8, 8, generator=g), "masks": None, "categorical": False},
             "mod_2": {"data": torch.rand(4, 5, 27, generator=g), "masks": None, "categorical": False}}
    enc = ours.encode(batch)
    ref_enc = ref_model.encode(batch)
    for name in FEATURE_DIMS:
        for part in ("shared", "private"):
            a, b = enc[name][part], ref_enc[name][part]
            assert (a is None) == (b is None)
            if a is not None:
                assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    missing, present = ours.get_missing_modalities({"mod_1": {"data": None}, "mod_2": {"data": 1}})
    assert missing == ["mod_1"] and present == ["mod_2"]
    # no CPU fallback: the objective needs the CUDA path
    with pytest.raises(RuntimeError):
        ours.objective(batch)


def test_registry_matches_reference_module_attributes():
    import mmvae_b200
    models, _, _ = ref_inplace.load()
    for name in ("poe", "moe", "mopoe", "dmvae"):
        assert hasattr(models, name) and hasattr(mmvae_b200, name)
        assert getattr(mmvae_b200, name).__name__ == getattr(models, name).__name__
