"""Drop-in for the reference's src/flux/condition.py (Condition, condition_dict; condition.py:10-138).

A Condition is built from a raw picture (PIL; pre-processed on the host like condition.py:53-90 and VAE-encoded when
the pipeline has a VAE attached), from already-encoded latents ([B, 16, h, w]) or from packed tokens ([B, N, 64]).  The position arithmetic on the
ids (condition.py:126-137) is kept operation for operation so the resulting ids are bit-identical.
"""
from typing import Tuple

import torch

from .pipeline_tools import encode_images

# type name -> type id (condition.py:10-21); the ids only matter as labels, `condition_type_ids` is ignored downstream
_TYPE_IDS = (("depth", 0), ("canny", 1), ("subject", 4), ("coloring", 6), ("deblurring", 7), ("depth_pred", 8), ("fill", 9),
             ("sr", 10), ("cartoon", 11), ("eeg+fnirs", 12))
condition_dict = dict(_TYPE_IDS)
_IMAGE_TYPES = tuple(name for name, _ in _TYPE_IDS if name != "eeg+fnirs")  # the types encode() accepts (condition.py:110-120)


def _rgb(img):
    return img.convert("RGB")


def _gray(img):  # "coloring": the luminance image, as three equal channels
    return img.convert("L").convert("RGB")


def _blurred(img):  # "deblurring": Gaussian blur of radius 10
    from PIL import ImageFilter

    return _rgb(_rgb(img).filter(ImageFilter.GaussianBlur(10)))


def _canny(img):  # "canny": OpenCV edges with thresholds (100, 200)
    import cv2
    import numpy as np
    from PIL import Image

    return _rgb(Image.fromarray(cv2.Canny(np.array(img), 100, 200)))


def _depth(img):
    raise NotImplementedError("'depth' runs a depth-estimation network (LiheYoung/depth-anything-small-hf through "
                              "transformers.pipeline, condition.py:59-68): not part of this build, pass the depth map as "
                              "`condition=` instead")


# condition type -> raw picture -> condition picture (condition.py:59-89)
_PREPROCESS = {"depth": _depth, "canny": _canny, "subject": lambda img: img, "coloring": _gray, "deblurring": _blurred,
               "fill": _rgb, "cartoon": _rgb}


def _shift_scale_ids(ids: torch.Tensor, delta, scale: float) -> torch.Tensor:
    """Position arithmetic of condition.py:126-137 on the (row, col) columns of `ids`, in the reference's operation
    order (add the delta, multiply by the scale, add (scale - 1) / 2) so the results are bit-identical."""
    rc = (1, 2)
    if delta is not None:
        for col, d in zip(rc, delta):
            ids[:, col] += d
    if scale != 1.0:
        half = (scale - 1.0) / 2
        for col in rc:
            ids[:, col] *= scale
        for col in rc:
            ids[:, col] += half
    return ids


class Condition(object):
    def __init__(self, condition_type: str, raw_img=None, condition=None, mask=None, position_delta=None,
                 position_scale=1.0, eeg=None, fnirs=None, ppg=None, motion=None) -> None:
        self.condition_type = condition_type
        assert raw_img is not None or condition is not None
        if raw_img is not None:
            self.condition = self.get_condition(condition_type, raw_img)
        else:
            self.condition = condition
        self.position_delta = position_delta
        self.position_scale = position_scale
        self.eeg, self.fnirs, self.ppg, self.motion = eeg, fnirs, ppg, motion
        assert mask is None, "Mask not supported yet"

    def get_condition(self, condition_type: str, raw_img):
        """condition.py:53-90: the condition image derived from the raw picture (host-side PIL / OpenCV work, once per
        edit).  Tensors (already-encoded latents or [0, 1] pictures) pass through untouched."""
        if isinstance(raw_img, torch.Tensor):
            return raw_img
        prepare = _PREPROCESS.get(condition_type)
        if prepare is None:  # the reference falls through to `self.condition`, which does not exist yet at this point
            raise NotImplementedError(f"no image preprocessing for condition type {condition_type!r}")
        return prepare(raw_img)

    @property
    def type_id(self) -> int:
        return condition_dict[self.condition_type]

    @classmethod
    def get_type_id(cls, condition_type: str) -> int:
        return condition_dict[condition_type]

    def encode(self, pipe) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """-> (tokens [B, N, 64], ids [N, 3], type_id [N, 1])."""
        if self.condition_type not in _IMAGE_TYPES:
            raise NotImplementedError(f"Condition type {self.condition_type} not implemented")
        c = self.condition
        if isinstance(c, torch.Tensor) and c.dim() == 3:  # packed tokens of a square latent grid
            tokens = c.to(pipe.device).to(pipe.dtype)
            side = int(round(tokens.shape[1] ** 0.5))
            assert side * side == tokens.shape[1], "packed condition tokens must come from a square latent grid"
            ids = pipe._prepare_latent_image_ids(tokens.shape[0], 2 * side, 2 * side, pipe.device, pipe.dtype)
        else:
            tokens, ids = encode_images(pipe, c)
        if self.position_delta is None and self.condition_type == "subject":
            width_px = c.size[0] if hasattr(c, "size") and not isinstance(c, torch.Tensor) else 16 * int(round(tokens.shape[1] ** 0.5))
            self.position_delta = [0, -width_px // 16]
        ids = _shift_scale_ids(ids, self.position_delta, self.position_scale)
        type_id = torch.ones_like(ids[:, :1]) * self.type_id
        return tokens, ids, type_id
