"""B200-native hot path of kywind/self-corr-pose (DINO features -> correspondence -> SoftRas -> losses)."""
__version__ = '0.1.0'
