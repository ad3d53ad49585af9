"""mmvae_b200 -- B200-native latent + objective hot path of multimodal-vae-comparison."""
__version__ = "0.1.0"
