"""diffusion_ccsp_b200 — B200-native reverse-diffusion CCSP sampling path."""
__version__ = "0.1.0"
