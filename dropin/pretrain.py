"""`from pretrain import Generator, MonetPhotoDataset, set_seed` drop-in (reference: pretrain.py:13-97); `pretrain_step` is the
body of the reference's training loop (pretrain.py:150-166) as a function."""
from multi_style_transfer_gan_b200.pretrain import Generator, MonetPhotoDataset, pretrain_step, set_seed  # noqa: F401
