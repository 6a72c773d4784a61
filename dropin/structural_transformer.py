"""The reference imports this module but does not ship it (enhanced_generator.py:4)."""
from multi_style_transfer_gan_b200.enhanced_generator import StructuralTransformerBlock  # noqa: F401
