"""Container-only: the drop-in classes must consume the RNG exactly like the reference's
constructors, so `torch.manual_seed(s); Model(...)` yields bit-identical parameters and buffers
(this is what lets goldens for the c=64 model be stored without their 10 MB of weights)."""
import pytest
import torch

pytestmark = pytest.mark.reference


@pytest.mark.parametrize("c,nb", [(16, 1), (64, 3), (8, 2)])
def test_generator_init_bit_identical(c, nb):
    from oracle import ref_import
    eg, _ = ref_import.load()
    from multi_style_transfer_gan_b200.enhanced_generator import EnhancedGenerator
    torch.manual_seed(5)
    ref = eg.EnhancedGenerator(channels=c, num_transformer_blocks=nb)
    torch.manual_seed(5)
    mine = EnhancedGenerator(channels=c, num_transformer_blocks=nb)
    rs, ms = ref.state_dict(), mine.state_dict()
    assert list(rs.keys()) == list(ms.keys())
    for k in rs:
        assert torch.equal(rs[k], ms[k]), k
    assert [n for n, _ in ref.named_children()] == [n for n, _ in mine.named_children()]
    assert [n for n, _ in ref.named_parameters()] == [n for n, _ in mine.named_parameters()]
    mine.load_state_dict(rs, strict=True)   # convert_model.py / pth_info.py contract: strict round trip
    ref.load_state_dict(ms, strict=True)


@pytest.mark.parametrize("c", [16, 8])
def test_discriminator_init_bit_identical(c):
    from oracle import ref_import
    eg, _ = ref_import.load()
    from multi_style_transfer_gan_b200.enhanced_generator import EnhancedDiscriminator
    torch.manual_seed(9)
    ref = eg.EnhancedDiscriminator(channels=c)
    torch.manual_seed(9)
    mine = EnhancedDiscriminator(channels=c)
    rs, ms = ref.state_dict(), mine.state_dict()
    assert list(rs.keys()) == list(ms.keys())
    for k in rs:
        assert torch.equal(rs[k], ms[k]), k
    assert [n for n, _ in ref.named_parameters()] == [n for n, _ in mine.named_parameters()]
    ref.load_state_dict(ms, strict=True)
    mine.load_state_dict(rs, strict=True)


def test_oracle_restatement_matches_live_reference():
    """restate.py vs the imported reference on fresh random weights (not only the goldens)."""
    from oracle import ref_import, restate as R
    from tests.util import assert_parity
    eg, _ = ref_import.load()
    torch.manual_seed(21)
    ref = eg.EnhancedGenerator(channels=8, num_transformer_blocks=2).eval()
    x = torch.rand(1, 3, 32, 48) * 2 - 1
    with torch.no_grad():
        assert_parity(R.generator_forward(dict(ref.state_dict()), x), ref(x), 1e-4, "G")
    D = eg.EnhancedDiscriminator(channels=8).eval()
    x = torch.rand(3, 3, 64, 64) * 2 - 1
    with torch.no_grad():
        s, st = D(x)
        s2, st2, _ = R.discriminator_forward(dict(D.state_dict()), x, training=False)
    assert_parity(s2, s, 1e-4, "score")
    assert_parity(st2, st, 1e-4, "struct")


@pytest.mark.parametrize("c", [64, 8])
def test_pretrain_generator_init_bit_identical(c):
    """pretrain.Generator (pretrain.py:60-97): same module tree, state_dict keys (incl. BatchNorm buffers) and default
    init in the same RNG order."""
    import importlib
    from oracle import ref_import
    ref_import.load()
    pt = importlib.import_module("pretrain")
    from multi_style_transfer_gan_b200.pretrain import Generator
    torch.manual_seed(3)
    ref = pt.Generator(channels=c)
    torch.manual_seed(3)
    mine = Generator(channels=c)
    rs, ms = ref.state_dict(), mine.state_dict()
    assert list(rs.keys()) == list(ms.keys())
    for k in rs:
        assert torch.equal(rs[k], ms[k]), k
    assert [n for n, _ in ref.named_parameters()] == [n for n, _ in mine.named_parameters()]
    mine.load_state_dict(rs, strict=True)
    ref.load_state_dict(ms, strict=True)
