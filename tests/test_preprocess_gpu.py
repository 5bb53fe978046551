"""GPU parity of the preprocessing kernels (C ABI groma_preprocess_image) against the Pillow-pinned oracle: bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hw", [(448, 448), (480, 640), (333, 500), (1080, 1920), (97, 61), (448, 900), (1200, 448), (14, 14),
                                (3000, 2000), (6720, 449)])
def test_resize_and_normalize_bit_exact(hw):
    from groma_b200 import ops as G
    from groma_b200.preprocess import _byte_table, IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD
    from oracle import preprocess as P
    rng = np.random.default_rng(hw[0] * 31 + hw[1])
    img = rng.integers(0, 256, size=(hw[0], hw[1], 3), dtype=np.uint8)
    if hw[0] <= 500:
        img[::2, ::3] = 255
        img[1::2, 1::3] = 0        # strong ringing -> exercises clip8 on both sides
    want_u8 = P.resize_bicubic_u8(img, 448, 448)
    want_f32 = P.preprocess_ref(img)
    lut_np = _byte_table(IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD, 1 / 255)
    assert np.array_equal(lut_np, P.normalize_lut())
    dev = torch.device("cuda:0")
    out_f32 = torch.empty((3, 448, 448), dtype=torch.float32, device=dev)
    out_u8 = torch.empty((448, 448, 3), dtype=torch.uint8, device=dev)
    G.preprocess_image(torch.from_numpy(img).to(dev), torch.from_numpy(lut_np).to(dev), 448, out_f32=out_f32, out_u8=out_u8)
    torch.cuda.synchronize()
    assert np.array_equal(out_u8.cpu().numpy(), want_u8)
    assert np.array_equal(out_f32.cpu().numpy(), want_f32)


def test_processor_matches_reference_call_sequence():
    """GromaImageProcessor(images) == PIL resize((448,448)) + HF BitImageProcessor(do_resize=False, do_center_crop=False)."""
    PIL = pytest.importorskip("PIL.Image")
    tr = pytest.importorskip("transformers")
    from groma_b200.preprocess import GromaImageProcessor
    rng = np.random.default_rng(5)
    imgs = [rng.integers(0, 256, size=s, dtype=np.uint8) for s in [(375, 500, 3), (640, 427, 3), (448, 448, 3)]]
    got = GromaImageProcessor().preprocess([PIL.fromarray(a, "RGB") for a in imgs], return_tensors="pt")["pixel_values"]
    assert got.shape == (3, 3, 448, 448) and got.is_cuda and got.dtype == torch.float32
    proc = tr.BitImageProcessor(do_resize=False, do_center_crop=False, image_mean=[0.485, 0.456, 0.406], image_std=[0.229, 0.224, 0.225])
    for i, a in enumerate(imgs):
        pil = PIL.fromarray(a, "RGB").resize((448, 448))
        want = proc.preprocess(pil, return_tensors="np")["pixel_values"][0]
        np.testing.assert_allclose(got[i].cpu().numpy(), want, rtol=0, atol=2.5e-7)


def test_rejects_unsupported_inputs():
    from groma_b200 import ops as G
    from groma_b200.lib import GromaError
    dev = torch.device("cuda:0")
    lut = torch.zeros((3, 256), device=dev)
    out = torch.empty((3, 448, 448), device=dev)
    with pytest.raises(GromaError):      # more than 15x downscale needs more than 64 taps
        G.preprocess_image(torch.zeros((7000, 10, 3), dtype=torch.uint8, device=dev), lut, 448, out_f32=out)
    with pytest.raises(ValueError):
        G.preprocess_image(torch.zeros((10, 10, 4), dtype=torch.uint8, device=dev), lut, 448, out_f32=out)
