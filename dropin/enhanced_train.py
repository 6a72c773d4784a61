"""`from enhanced_train import EnhancedCycleGAN` drop-in (reference: enhanced_train.py:13-152)."""
from multi_style_transfer_gan_b200.enhanced_train import EnhancedCycleGAN  # noqa: F401
