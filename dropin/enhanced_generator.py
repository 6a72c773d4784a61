"""Put this directory first on sys.path and the reference's scripts
(`from enhanced_generator import EnhancedGenerator, EnhancedDiscriminator`) pick up the B200 path."""
from multi_style_transfer_gan_b200.enhanced_generator import (  # noqa: F401
    EnhancedDiscriminator, EnhancedGenerator, LocalAttention, MultiScaleBlock, StructuralTransformerBlock)
