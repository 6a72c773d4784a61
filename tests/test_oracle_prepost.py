"""The oracle's restatement of the reference's uint8 pre / post-processing (batch_process_images.py:193-205, 287-291, 304-310)
against the reference's OWN lines executed with PIL / torchvision / numpy (all present in the container): bit-exact."""
import numpy as np
import pytest
import torch

from oracle import restate as R


def _reference_lines(resized_np, target, strength, styled_np):
    """batch_process_images.py:193-205 and :304-310, verbatim semantics (PIL canvas paste, torchvision transforms, numpy blend)"""
    from PIL import Image
    from torchvision import transforms
    resized = Image.fromarray(resized_np)
    new_width, new_height = resized.size
    canvas = Image.new("RGB", target, (255, 255, 255))
    offset_x = (target[0] - new_width) // 2
    offset_y = (target[1] - new_height) // 2
    canvas.paste(resized, (offset_x, offset_y))
    transform = transforms.Compose([transforms.ToTensor(), transforms.Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))])
    x = transform(canvas)
    original_np = np.array(canvas)
    result_np = original_np * (1 - strength) + styled_np * strength
    result_np = np.clip(result_np, 0, 255).astype(np.uint8)
    return x, original_np, result_np, (offset_y, offset_x)


@pytest.mark.parametrize("hw", [(256, 171), (144, 256), (256, 256)])
def test_letterbox_normalize_and_strength_blend_match_the_reference_lines(hw):
    pytest.importorskip("PIL")
    pytest.importorskip("torchvision")
    rng = np.random.default_rng(hw[0] * 1000 + hw[1])
    resized = rng.integers(0, 256, size=hw + (3,), dtype=np.uint8)
    styled = rng.integers(0, 256, size=(256, 256, 3), dtype=np.uint8)
    for strength in (0.0, 0.35, 0.8, 1.0):
        x_ref, canvas_ref, blend_ref, (oy, ox) = _reference_lines(resized, (256, 256), strength, styled)
        x, canvas = R.letterbox_normalize(torch.from_numpy(resized), 256, 256, oy, ox)
        assert torch.equal(x, x_ref)
        assert np.array_equal(canvas.numpy(), canvas_ref)
        assert np.array_equal(R.strength_blend_u8(canvas, torch.from_numpy(styled), strength).numpy(), blend_ref)
