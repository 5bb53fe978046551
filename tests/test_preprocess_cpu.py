"""Pin the preprocessing oracle (oracle/preprocess.py) against the Pillow / transformers installed in this image."""
import numpy as np
import pytest

from oracle import preprocess as P

PIL = pytest.importorskip("PIL.Image")


@pytest.mark.parametrize("hw", [(448, 448), (480, 640), (333, 500), (1080, 1920), (97, 61), (448, 900), (1200, 448), (14, 14),
                                (2000, 3000)])
def test_resize_matches_pillow_bit_exact(hw):
    rng = np.random.default_rng(hw[0] * 7919 + hw[1])
    img = rng.integers(0, 256, size=(hw[0], hw[1], 3), dtype=np.uint8)
    want = np.asarray(PIL.fromarray(img, "RGB").resize((448, 448)))
    got = P.resize_bicubic_u8(img, 448, 448)
    assert got.dtype == np.uint8 and got.shape == (448, 448, 3)
    assert np.array_equal(got, want)


def test_resize_extreme_values_clip():
    img = np.zeros((200, 300, 3), dtype=np.uint8)
    img[::2, ::3] = 255            # ringing of the bicubic kernel must clip to [0, 255] exactly like clip8()
    want = np.asarray(PIL.fromarray(img, "RGB").resize((448, 448)))
    assert np.array_equal(P.resize_bicubic_u8(img, 448, 448), want)


def test_normalize_matches_transformers_processor():
    tr = pytest.importorskip("transformers")
    proc = tr.BitImageProcessor(do_resize=False, do_center_crop=False, do_rescale=True, rescale_factor=1 / 255, do_normalize=True,
                                image_mean=list(P.IMAGENET_MEAN), image_std=list(P.IMAGENET_STD), do_convert_rgb=True)
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, size=(448, 448, 3), dtype=np.uint8)
    img[0, :256, 0] = np.arange(256)     # every byte value at least once in channel 0
    want = proc.preprocess(PIL.fromarray(img, "RGB"), return_tensors="np")["pixel_values"][0]
    got = P.preprocess_ref(img)
    assert got.shape == want.shape == (3, 448, 448)
    # the installed transformers (5.x) may fuse rescale+normalize; 4.32 semantics agree to 1 float32 ulp
    np.testing.assert_allclose(got, want, rtol=0, atol=2.5e-7)


def test_full_pipeline_against_reference_call_sequence():
    tr = pytest.importorskip("transformers")
    proc = tr.BitImageProcessor(do_resize=False, do_center_crop=False, image_mean=list(P.IMAGENET_MEAN), image_std=list(P.IMAGENET_STD))
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, size=(375, 500, 3), dtype=np.uint8)
    pil = PIL.fromarray(img, "RGB").resize((448, 448))                      # run_groma.py:78
    want = proc.preprocess(pil, return_tensors="np")["pixel_values"][0]   # run_groma.py:79
    np.testing.assert_allclose(P.preprocess_ref(img), want, rtol=0, atol=2.5e-7)


def test_processor_host_logic_without_gpu():
    """Host side of GromaImageProcessor: the byte table equals the oracle's, inputs are validated, PIL images are converted."""
    import torch
    from groma_b200.preprocess import GromaImageProcessor, _byte_table, IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD
    assert np.array_equal(_byte_table(IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD, 1 / 255), P.normalize_lut())
    t = GromaImageProcessor._to_hwc_u8(PIL.fromarray(np.zeros((5, 7, 3), dtype=np.uint8), "RGB"))
    assert t.dtype == torch.uint8 and tuple(t.shape) == (5, 7, 3)
    t = GromaImageProcessor._to_hwc_u8(PIL.fromarray(np.zeros((5, 7), dtype=np.uint8), "L"))      # .convert('RGB') as the reference does
    assert tuple(t.shape) == (5, 7, 3)
    with pytest.raises(ValueError):
        GromaImageProcessor._to_hwc_u8(np.zeros((5, 7, 3), dtype=np.float32))
    with pytest.raises(ValueError):
        GromaImageProcessor().preprocess([np.zeros((5, 7, 3), dtype=np.uint8)], return_tensors="np")
