"""The reference's caller contract (inference.py:24-117): `load_model` builds OminiModel on the CPU in float32, loads a
LoRA file or a full state dict, then `model.to("cuda")` / `model.flux_pipe.to("cuda")`; `inference_single_image` drives
`generate()` with a PIL condition picture, a text prompt and raw signals.

Two legs, because the reference tree exists only in the build container and the GPU only on the GPU box:
  * not-gpu: the reference's OWN inference.py is imported (this repo's `src` first on sys.path) and `load_model` runs
    unchanged up to `model.to("cuda")`, which must raise the explicit "no CUDA device" error here (no CPU fallback);
    with a GPU present the same test runs `inference_single_image` too;
  * gpu: the same call sequence, statement for statement (line numbers cited), on the native engine.
"""
import importlib.util
import os
import sys
import types

import pytest
import torch

REF = "/root/reference/inference.py"


def _tiny_cfg():
    from loongx_b200.config import FluxConfig

    return FluxConfig(num_layers=1, num_single_layers=1, num_attention_heads=2, joint_attention_dim=4096,
                      pooled_projection_dim=768)


def _config(flux_path):
    """train/config/seed_512.yaml as inference.py reads it (dtype "float32", lora r = 4, the `model` block)."""
    return {"flux_path": flux_path, "dtype": "float32",
            "model": {"union_cond_attn": True, "add_cond_attn": False, "latent_lora": False},
            "train": {"lora_config": {"r": 4, "lora_alpha": 4, "init_lora_weights": "gaussian"}}}


def _import_reference_inference(monkeypatch):
    """/root/reference/inference.py as a module; `accelerate` (imported at its top, unused on this path) is stubbed for
    the duration of the test (a stand-in without a spec left in sys.modules would break libraries that probe for it)."""
    if "accelerate" not in sys.modules:
        try:
            import accelerate  # noqa: F401
        except ImportError:
            stub = types.ModuleType("accelerate")
            stub.init_empty_weights = stub.infer_auto_device_map = lambda *a, **k: None
            monkeypatch.setitem(sys.modules, "accelerate", stub)
    spec = importlib.util.spec_from_file_location("ref_inference", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _attach_small_encoders(model):
    """What FluxPipeline.from_pretrained provides from a checkpoint directory: VAE + both text encoders (synthetic)."""
    from loongx_b200.text import ClipTextConfig, NativeClipText, NativeT5Encoder, T5Config
    from oracle import text_encoders as OT

    pipe = model.flux_pipe
    pipe.attach_vae(None)
    ck = dict(vocab_size=99, hidden_size=768, intermediate_size=256, num_layers=1, num_heads=12, max_positions=77)
    tk = dict(vocab_size=99, d_model=4096, d_kv=64, d_ff=256, num_layers=1, num_heads=2)
    PC = {k: v.to(torch.bfloat16).float() for k, v in OT.clip_init(OT.ClipCfg(**ck), 5).items()}
    PT = {k: v.to(torch.bfloat16).float() for k, v in OT.t5_init(OT.T5Cfg(**tk), 6).items()}

    def fake_tokenizer(prompts, padding, max_length, truncation, return_tensors, **kw):
        g = torch.Generator().manual_seed(len(prompts[0]))
        ids = torch.randint(3, 98, (len(prompts), max_length), generator=g)
        ids[:, -1] = 98  # EOS = the largest id (CLIP pooling position)
        return {"input_ids": ids}

    pipe.attach_text_encoders(clip=NativeClipText(ClipTextConfig(**ck), PC, "cuda"),
                              t5=NativeT5Encoder(T5Config(**tk), PT, "cuda"), tokenizers=(fake_tokenizer, fake_tokenizer))


def _signals():
    g = torch.Generator().manual_seed(45)
    return dict(eeg=torch.randn(4, 3000, generator=g).numpy(), fnirs=torch.randn(6, 700, generator=g).numpy(),
                ppg=torch.randn(4, 256, generator=g).numpy(), motion=torch.randn(6, 100, generator=g).numpy())


@pytest.mark.skipif(not os.path.exists(REF), reason="the reference tree exists only in the build container")
def test_reference_inference_py_runs_against_this_repo(tmp_path, monkeypatch):
    from src.train import model as our_model

    ref = _import_reference_inference(monkeypatch)
    assert ref.OminiModel is our_model.OminiModel, "inference.py must bind this repo's src.train.model"
    # inference.py passes config["flux_path"] (a string) as flux_pipe_id: hand it a name that resolves to a tiny config
    real_init = our_model.OminiModel.__init__

    def init(self, flux_pipe_id, *a, **k):
        real_init(self, _tiny_cfg() if flux_pipe_id == "tiny-synthetic" else flux_pipe_id, *a, **k)

    monkeypatch.setattr(our_model.OminiModel, "__init__", init)
    lora_dir = tmp_path / "lora_ckpt"  # "lora" in the path -> the load_lora branch (inference.py:43-44)
    lora_dir.mkdir()
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CUDA device"):
            ref.load_model(str(lora_dir), config=_config("tiny-synthetic"))
        return
    # with a GPU: save a LoRA file first, then the whole reference flow
    seed_model = our_model.OminiModel("tiny-synthetic", lora_config={"r": 4, "lora_alpha": 4}, device="cuda")
    seed_model.save_lora(str(lora_dir))
    model = ref.load_model(str(lora_dir), config=_config("tiny-synthetic"))
    assert model.device.type == "cuda" and not model.training
    _attach_small_encoders(model)
    from PIL import Image

    # condition_type: inference.py's default "SEED" is not a type Condition.encode accepts (condition.py:110-124 raises
    # NotImplementedError in the reference too); its __main__ passes the config's "subject"
    img = ref.inference_single_image(model, Image.new("RGB", (64, 64), (120, 30, 200)), "a cat", condition_type="subject",
                                     target_size=64, **{k + "_data": v for k, v in _signals().items()})
    assert img.size == (64, 64)


def test_staged_model_records_checkpoint_loads_and_refuses_cpu_compute():
    """OminiModel(device="cpu", dtype=float32) (inference.py:35-41) builds nothing native; load_lora / load_state_dict
    are replayed by .to("cuda"); any use before that raises."""
    from src.train.model import OminiModel

    m = OminiModel(_tiny_cfg(), lora_config={"r": 4, "lora_alpha": 4}, device="cpu", dtype=torch.float32,
                   model_config={"union_cond_attn": True})
    assert m.staged and m.device.type == "cpu" and m._dtype == torch.float32
    assert m.load_lora("/some/lora/dir") is m and m._pending == [("lora", "/some/lora/dir")]
    sd = {k: v for k, v in torch.nn.Module.state_dict(m).items()}
    sd["transformer.x_embedder.weight"] = torch.zeros(3, 3)
    m.load_state_dict(sd)
    assert m._pending[-1][0] == "state" and "x_embedder.weight" in m._pending[-1][1]
    with pytest.raises(RuntimeError, match="staged"):
        m.flux_pipe.vae
    assert m.to("cpu") is m and m.to(torch.float32) is m
    with pytest.raises(NotImplementedError):
        m.to(torch.float16)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CUDA device"):
            m.to("cuda")
    with pytest.raises(NotImplementedError):
        OminiModel(_tiny_cfg(), lora_config={"r": 4}, device="cpu", dtype=torch.float16)


@pytest.mark.gpu
def test_inference_py_call_sequence_on_the_native_engine(tmp_path):
    """load_model (inference.py:24-60) + inference_single_image (:77-117), statement for statement, for both checkpoint
    kinds; the staged model must produce exactly what a model built directly on the GPU produces."""
    from PIL import Image

    from src.flux.condition import Condition
    from src.flux.generate import generate
    from src.train.model import OminiModel

    config = _config(_tiny_cfg())
    src_model = OminiModel(_tiny_cfg(), lora_config=config["train"]["lora_config"], device="cuda")
    with torch.no_grad():
        for p in src_model.lora_layers:  # non-trivial LoRA factors
            p.normal_(0, 0.05)
    src_model.transformer.remerge_lora()
    lora_dir = tmp_path / "lora_ckpt"
    src_model.save_lora(str(lora_dir))
    full_path = tmp_path / "full.ckpt"
    torch.save({"state_dict": {k: v.detach().cpu() for k, v in src_model.state_dict().items()}}, str(full_path))

    def load_model(checkpoint_path):
        model = OminiModel(flux_pipe_id=config["flux_path"], lora_config=config["train"]["lora_config"], device="cpu",
                           dtype=getattr(torch, config["dtype"]), model_config=config.get("model", {}))  # :35-41
        if "lora" in checkpoint_path:  # :43-44
            model.load_lora(checkpoint_path)
        else:  # :45-53
            checkpoint = torch.load(checkpoint_path, map_location=model.device)
            model.load_state_dict(checkpoint["state_dict"] if "state_dict" in checkpoint else checkpoint)
        model.to("cuda")  # :55
        model.flux_pipe.to("cuda")  # :56
        model.eval()  # :58
        return model

    def inference_single_image(model, condition_img, prompt, seed=42, **sig):  # :77-117
        generator = torch.Generator(device=model.device)
        generator.manual_seed(seed)
        condition = Condition(condition_type="subject", condition=condition_img, position_delta=[0, 0], eeg=sig["eeg"],
                              fnirs=sig["fnirs"], ppg=sig["ppg"], motion=sig["motion"])
        result = generate(model, model.flux_pipe, prompt=prompt, conditions=[condition], height=64, width=64,
                          generator=generator, model_config=model.model_config, default_lora=True,
                          additional_condition1=sig["eeg"], additional_condition2=sig["fnirs"],
                          additional_condition3=sig["ppg"], additional_condition4=sig["motion"], use_brain_condition=True,
                          fuse_flag=False, num_inference_steps=2)
        return result.images[0]

    img_in = Image.new("RGB", (64, 64), (120, 30, 200))
    outs = []
    for ckpt in (str(lora_dir), str(full_path)):
        model = load_model(ckpt)
        assert model.device.type == "cuda" and not model.staged and not model.training
        if ckpt == str(lora_dir):
            # a LoRA-only checkpoint carries no CS3 / DGF weights: take them from the source model for the comparison
            torch.nn.Module.load_state_dict(model, torch.nn.Module.state_dict(src_model))
            # (the base DiT weights come from the constructor's default seed, like src_model's)
        _attach_small_encoders(model)
        torch.manual_seed(123)  # encode_images draws latent_dist.sample() from the global generator (pipeline_tools.py:10)
        outs.append(inference_single_image(model, img_in, "a cat", **_signals()))
    _attach_small_encoders(src_model)
    src_model.eval()
    torch.manual_seed(123)
    ref_img = inference_single_image(src_model, img_in, "a cat", **_signals())
    import numpy as np

    # LoRA file: factors re-merged by the same kernel -> bit-identical picture.  Full state dict: the panels are re-packed
    # from the exported tensors (W + s B A summed on the host path of pack_linear), which may differ from the merge kernel in
    # the last bf16 bit of a few weights -> at most 2/255 on any pixel.
    ref = np.asarray(ref_img).astype(np.int32)
    diffs = [int(np.abs(np.asarray(o).astype(np.int32) - ref).max()) for o in outs]
    print(f"\n[inference.py sequence] max |pixel difference| vs the source model: LoRA checkpoint {diffs[0]}, full checkpoint {diffs[1]}")
    assert len(outs) == 2 and diffs[0] == 0 and diffs[1] <= 2, diffs
