"""not-gpu: the host restatement of Pillow's uint8 BICUBIC resample is bit-identical to PIL itself
(the parent -> condition resize of tts/tts_reflectionflow.py:276-277)."""
import numpy as np
import pytest
from PIL import Image

from reflectionflow_b200.resize import precompute_coeffs, resize_u8_reference


@pytest.mark.parametrize("hw,out", [((64, 64), (32, 32)), ((96, 128), (48, 64)), ((100, 60), (37, 45)),
                                     ((32, 32), (64, 64)), ((256, 256), (128, 128))])
def test_matches_pil_bit_exact(hw, out):
    rng = np.random.default_rng(hw[0] * 1000 + out[0])
    img = rng.integers(0, 256, size=(hw[0], hw[1], 3), dtype=np.uint8)
    want = np.array(Image.fromarray(img).resize((out[1], out[0])))  # PIL default = BICUBIC
    got = resize_u8_reference(img, out[0], out[1])
    assert np.array_equal(got, want)


def test_coeff_table_shape_for_the_2x_downscale():
    b, k = precompute_coeffs(1024, 512)
    assert k.shape == (512, 9) and b.shape == (512, 2)
    assert b[100].tolist() == [197, 8]  # interior: 8 taps
    assert int(k[100].sum()) in range((1 << 22) - 8, (1 << 22) + 9)
