"""`from pretrain import Generator` drop-in (reference: pretrain.py:60-97); `pretrain_step` is the body of the
reference's training loop (pretrain.py:150-166) as a function."""
from multi_style_transfer_gan_b200.pretrain import Generator, pretrain_step  # noqa: F401
