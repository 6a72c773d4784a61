"""The reference's own import lines resolve against dropin/ (the shim directory a maintainer puts first on sys.path,
INTEGRATION.md).  CPU only: construction and state_dict layout, no kernels."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2])
# enhanced_train.py:10-11, batch_process_images.py:18, m_test.py:12, enhanced_generator.py:4 of the reference
from enhanced_generator import EnhancedGenerator, EnhancedDiscriminator
from pretrain import MonetPhotoDataset, set_seed
from pretrain import Generator
from structural_transformer import StructuralTransformerBlock
from enhanced_train import EnhancedCycleGAN
import enhanced_generator, multi_style_transfer_gan_b200.enhanced_generator as ours
assert enhanced_generator.EnhancedGenerator is ours.EnhancedGenerator
set_seed(1)
g = EnhancedGenerator(16, 1)
keys = list(g.state_dict().keys())
assert keys[0] == 'initial.0.weight' and g.state_dict()[keys[0]].shape[0] == 16     # direct_transform.py:25-28 reads this
assert [n for n, _ in g.named_children()] == ['initial', 'down1', 'down2', 'transformer_blocks', 'up1', 'up2', 'output', 'style_encoder']
d = EnhancedDiscriminator(16)
assert len(d.state_dict()) == 28
import tempfile, os
root = tempfile.mkdtemp(); os.makedirs(os.path.join(root, 'trainA'))
assert len(MonetPhotoDataset(root, 'A')) == 0
print('dropin ok')
"""


def test_reference_import_lines_resolve_against_dropin():
    r = subprocess.run([sys.executable, "-c", SCRIPT, ROOT, os.path.join(ROOT, "dropin")], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "dropin ok" in r.stdout
