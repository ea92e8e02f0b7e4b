"""Host-side condition preprocessing (reference condition.py:53-90) and the image processor's PIL path."""
import numpy as np
import pytest
import torch

PIL = pytest.importorskip("PIL.Image")


def _picture(w=48, h=32):
    rng = np.random.default_rng(0)
    return PIL.fromarray(rng.integers(0, 255, (h, w, 3), dtype=np.uint8))


def test_condition_pictures_by_type():
    from PIL import ImageFilter

    from src.flux.condition import Condition, condition_dict

    img = _picture()
    same = lambda a, b: np.array_equal(np.asarray(a), np.asarray(b))  # noqa: E731
    assert Condition("subject", raw_img=img).condition is img
    assert same(Condition("fill", raw_img=img.convert("RGBA")).condition, img)
    assert same(Condition("cartoon", raw_img=img).condition, img)
    gray = Condition("coloring", raw_img=img).condition
    assert gray.mode == "RGB" and same(gray, img.convert("L").convert("RGB"))
    assert same(Condition("deblurring", raw_img=img).condition, img.filter(ImageFilter.GaussianBlur(10)))
    cv2 = pytest.importorskip("cv2")
    edges = Condition("canny", raw_img=img).condition
    assert edges.mode == "RGB" and same(np.asarray(edges)[..., 0], cv2.Canny(np.asarray(img), 100, 200))
    with pytest.raises(NotImplementedError, match="depth"):
        Condition("depth", raw_img=img)
    with pytest.raises(NotImplementedError):
        Condition("sr", raw_img=img)
    t = torch.zeros(1, 16, 4, 4)
    assert Condition("depth", raw_img=t).condition is t  # tensors pass through
    assert Condition.get_type_id("canny") == condition_dict["canny"] == 1


def test_image_processor_resizes_to_the_vae_grid():
    from loongx_b200.vae import ImageProcessor

    ip = ImageProcessor(16)
    x = ip.preprocess(_picture(50, 37))  # -> 48 x 32, lanczos, like VaeImageProcessor(do_resize=True)
    assert x.shape == (1, 3, 32, 48) and x.min() >= -1 and x.max() <= 1
    want = np.asarray(_picture(50, 37).resize((48, 32), resample=PIL.LANCZOS), dtype=np.float32) / 255.0
    assert torch.allclose(x[0].permute(1, 2, 0), torch.from_numpy(2 * want - 1), atol=1e-6)
    with pytest.raises(ValueError):
        ip.preprocess(_picture(8, 40))
    with pytest.raises(ValueError):
        ip.preprocess([_picture(48, 32), _picture(32, 32)])
