"""Drop-in for the reference's src/train/data.py::SeedDataset (data.py:11-98): the record reader of the LoongX training /
test sets.

`<dir>/<name>.jsonl` holds one JSON object per line (`source_image`, `target_image`, `speech2text` or `instruction`);
`<dir>/data_final.pkl` is a pickled dict `file name -> {"EEG": [4, L], "FNIRS": [6, L], "PPG": [4, L], "Motion": [6, L]}`
(variable L).  Records whose source image has no signals are dropped (data.py:46-50).  `__getitem__` returns the dict the
reference's `OminiModel.step` consumes (data.py:87-98).

The VAE and the text encoders are outside this build, so two optional sidecars make a record
directly consumable by the native `OminiModel.step`:
  * `latent_dir`: `<latent_dir>/<image file name>.pt` = pre-encoded, shifted / scaled VAE latents [16, h, w];
    when present, `image` / `condition` are those tensors instead of pixels;
  * `embed_dir`: `<embed_dir>/<image file name>.pt` = {"prompt_embeds": [N, 4096], "pooled_prompt_embeds": [768]}.
Host-side I/O only.
"""
import json
import os
import pickle

import numpy as np
import torch
from torch.utils.data import Dataset


class SeedDataset(Dataset):
    def __init__(self, jsonl_path, condition_size: int = 512, condition_type: str = "subject", image_dir="", transform=None,
                 return_pil_image=False, latent_dir: str = None, embed_dir: str = None):
        self.samples = []
        self.image_dir = image_dir
        self.transform = transform
        self.return_pil_image = return_pil_image
        self.condition_type = condition_type
        self.condition_size = condition_size
        self.latent_dir, self.embed_dir = latent_dir, embed_dir
        pkl_path = os.path.join(os.path.dirname(jsonl_path), "data_final.pkl")
        with open(pkl_path, "rb") as f:
            self.bio_data = pickle.load(f)
        with open(jsonl_path, "r", encoding="utf-8") as f:
            for line in f:
                if not line.strip():
                    continue
                aline = json.loads(line)
                if aline["source_image"].split("/")[-1] in self.bio_data:
                    self.samples.append(aline)

    def __len__(self):
        return len(self.samples)

    def _image(self, rel_path):
        name = rel_path.split("/")[-1]
        if self.latent_dir is not None:
            return torch.load(os.path.join(self.latent_dir, name + ".pt"), map_location="cpu", weights_only=True)
        from PIL import Image

        img = Image.open(os.path.join(self.image_dir, rel_path)).convert("RGB")
        if self.return_pil_image:
            return img
        if self.transform is not None:
            return self.transform(img)
        img = img.resize((512, 512), Image.BILINEAR)  # T.Resize((512, 512)) + T.ToTensor() (data.py:53-57)
        return torch.from_numpy(np.asarray(img, dtype=np.float32) / 255.0).permute(2, 0, 1).contiguous()

    def __getitem__(self, idx):
        item = self.samples[idx]
        name = item["source_image"].split("/")[-1]
        bio = self.bio_data[name]
        out = {
            "image": self._image(item["source_image"]),
            "condition": self._image(item["target_image"]),
            "description": item["speech2text"] if "speech2text" in item else item["instruction"],
            "condition_type": self.condition_type,
            "position_delta": np.array([0, -self.condition_size // 16]),
            "eeg": np.array(bio["EEG"]),
            "fnirs": np.array(bio["FNIRS"]) if "FNIRS" in bio else None,
            "ppg": np.array(bio["PPG"]) if "PPG" in bio else None,
            "motion": np.array(bio["Motion"]) if "Motion" in bio else None,
        }
        if self.embed_dir is not None:
            out.update(torch.load(os.path.join(self.embed_dir, name + ".pt"), map_location="cpu", weights_only=True))
        return out


def collate_step_batch(records):
    """Batch of SeedDataset records -> the dict `OminiModel.step` takes: tensors stacked, raw signals zero-padded to the
    longest record (the model's length normaliser pads / truncates to its fixed lengths anyway, model.py:479-511)."""
    def stack_signal(key):
        xs = [r[key] for r in records]
        if any(x is None for x in xs):
            return None
        L = max(x.shape[-1] for x in xs)
        out = torch.zeros(len(xs), xs[0].shape[0], L, dtype=torch.float32)
        for i, x in enumerate(xs):
            out[i, :, :x.shape[-1]] = torch.as_tensor(np.asarray(x), dtype=torch.float32)
        return out

    batch = {
        "image": torch.stack([torch.as_tensor(r["image"]) for r in records]),
        "condition": torch.stack([torch.as_tensor(r["condition"]) for r in records]),
        "description": [r["description"] for r in records],
        "condition_type": [r["condition_type"] for r in records],
        "position_delta": [records[0]["position_delta"].tolist()],
    }
    for key in ("eeg", "fnirs", "ppg", "motion"):
        v = stack_signal(key)
        if v is not None:
            batch[key] = v
    if "prompt_embeds" in records[0]:
        batch["prompt_embeds"] = torch.stack([r["prompt_embeds"] for r in records])
        batch["pooled_prompt_embeds"] = torch.stack([r["pooled_prompt_embeds"] for r in records])
    return batch
