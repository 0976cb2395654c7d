"""nmf_b200 -- B200-native per-ray rendering hot path of Neural Microfacet Fields."""
__version__ = "0.1.0"
