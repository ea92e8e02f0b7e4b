"""CPU test of the SeedDataset record reader (reference data.py:11-98) on a synthetic data set."""
import json
import pickle

import numpy as np
import torch


def test_seed_dataset_records_and_collation(tmp_path):
    from PIL import Image

    from src.train.data import SeedDataset, collate_step_batch

    rng = np.random.default_rng(0)
    names = ["a.png", "b.png", "c.png"]
    for n in names + ["a_t.png", "b_t.png", "c_t.png"]:
        Image.fromarray(rng.integers(0, 255, (40, 56, 3), dtype=np.uint8)).save(tmp_path / n)
    bio = {"a.png": {"EEG": rng.normal(size=(4, 5000)).tolist(), "FNIRS": rng.normal(size=(6, 600)).tolist(),
                     "PPG": rng.normal(size=(4, 256)).tolist(), "Motion": rng.normal(size=(6, 100)).tolist()},
           "b.png": {"EEG": rng.normal(size=(4, 3000)).tolist()}}  # c.png has no signals -> dropped (data.py:46-50)
    with open(tmp_path / "data_final.pkl", "wb") as f:
        pickle.dump(bio, f)
    with open(tmp_path / "testset.jsonl", "w") as f:
        f.write(json.dumps({"source_image": "imgs/a.png", "target_image": "a_t.png", "instruction": "make it red"}) + "\n")
        f.write(json.dumps({"source_image": "b.png", "target_image": "b_t.png", "instruction": "x", "speech2text": "blue"}) + "\n")
        f.write(json.dumps({"source_image": "c.png", "target_image": "c_t.png", "instruction": "y"}) + "\n")
    (tmp_path / "imgs").mkdir()
    (tmp_path / "a.png").rename(tmp_path / "imgs" / "a.png")
    ds = SeedDataset(str(tmp_path / "testset.jsonl"), image_dir=str(tmp_path))
    assert len(ds) == 2
    r0, r1 = ds[0], ds[1]
    assert r0["image"].shape == (3, 512, 512) and 0.0 <= float(r0["image"].min()) and float(r0["image"].max()) <= 1.0
    assert r0["description"] == "make it red" and r1["description"] == "blue"  # speech2text wins (data.py:90)
    assert r0["eeg"].shape == (4, 5000) and r0["motion"].shape == (6, 100) and r1["fnirs"] is None
    assert r0["position_delta"].tolist() == [0, -32] and r0["condition_type"] == "subject"
    # sidecars: pre-encoded latents + text embeddings make the record consumable by the native OminiModel.step
    lat, emb = tmp_path / "lat", tmp_path / "emb"
    lat.mkdir()
    emb.mkdir()
    for n in ("a.png", "a_t.png"):
        torch.save(torch.randn(16, 64, 64), lat / (n + ".pt"))
    torch.save({"prompt_embeds": torch.randn(512, 4096), "pooled_prompt_embeds": torch.randn(768)}, emb / "a.png.pt")
    ds2 = SeedDataset(str(tmp_path / "testset.jsonl"), latent_dir=str(lat), embed_dir=str(emb))
    b = collate_step_batch([ds2[0]])
    assert b["image"].shape == (1, 16, 64, 64) and b["prompt_embeds"].shape == (1, 512, 4096)
    assert b["eeg"].shape == (1, 4, 5000) and b["position_delta"] == [[0, -32]] and b["condition_type"] == ["subject"]
    # ragged signals are zero-padded to the longest record
    b2 = collate_step_batch([{**ds[0], "image": torch.zeros(3, 8, 8), "condition": torch.zeros(3, 8, 8)},
                             {**ds[1], "image": torch.zeros(3, 8, 8), "condition": torch.zeros(3, 8, 8)}])
    assert b2["eeg"].shape == (2, 4, 5000) and float(b2["eeg"][1, :, 3000:].abs().max()) == 0.0 and "fnirs" not in b2
