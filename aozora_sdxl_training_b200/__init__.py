"""aozora-b200: B200-native SDXL UNet training step (drop-in for the hot path of Hysocs/Aozora_SDXL_Training)."""
__version__ = "0.1.0"
