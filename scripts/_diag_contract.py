import os, sys
sys.path.insert(0, os.getcwd())
sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, numpy as np
import test_inference_contract as T
from src.train.model import OminiModel
from src.flux.condition import Condition
from src.flux.generate import generate
from PIL import Image
import tempfile
tmp = tempfile.mkdtemp()
config = T._config(T._tiny_cfg())
src = OminiModel(T._tiny_cfg(), lora_config=config["train"]["lora_config"], device="cuda")
with torch.no_grad():
    for p in src.lora_layers: p.normal_(0, 0.05)
src.transformer.remerge_lora()
src.save_lora(tmp + "/lora_ckpt")
m = OminiModel(flux_pipe_id=config["flux_path"], lora_config=config["train"]["lora_config"], device="cpu", dtype=torch.float32, model_config=config.get("model", {}))
m.load_lora(tmp + "/lora_ckpt"); m.to("cuda"); m.flux_pipe.to("cuda"); m.eval()
torch.nn.Module.load_state_dict(m, torch.nn.Module.state_dict(src))
src.eval()
# weights
Ps, Pm = src.transformer.weights.export_params(), m.transformer.weights.export_params()
bad = [k for k in Ps if not torch.equal(Ps[k].float(), Pm[k].float())]
print("exported params differing:", len(bad), bad[:5])
for key, pa in src.transformer.weights.named.items():
    pb = m.transformer.weights.named[key]
    if hasattr(pa, "w_lora") and pa.w_lora is not None and not torch.equal(pa.w_lora, pb.w_lora):
        print("merged panel differs:", key, (pa.w_lora.float() - pb.w_lora.float()).abs().max().item())
for (n1, p1), (n2, p2) in zip(src.named_parameters(), m.named_parameters()):
    if not torch.equal(p1, p2): print("cs3 param differs", n1)
for mm in (src, m): T._attach_small_encoders(mm)
sig = T._signals()
img = Image.new("RGB", (64, 64), (120, 30, 200))
def run(model, out="latent"):
    torch.manual_seed(123)
    gen = torch.Generator(device=model.device); gen.manual_seed(42)
    c = Condition(condition_type="subject", condition=img, position_delta=[0, 0])
    return generate(model, model.flux_pipe, prompt="a cat", conditions=[c], height=64, width=64, generator=gen, model_config=model.model_config,
                    default_lora=True, additional_condition1=sig["eeg"], additional_condition2=sig["fnirs"], additional_condition3=sig["ppg"],
                    additional_condition4=sig["motion"], use_brain_condition=True, fuse_flag=False, num_inference_steps=2, output_type=out).images
a1, a2, b1 = run(src), run(src), run(m)
print("src twice equal:", torch.equal(a1, a2), " src vs loaded equal:", torch.equal(a1, b1), (a1.float() - b1.float()).abs().max().item())
print("model_config", src.model_config, m.model_config)
from src.flux.pipeline_tools import prepare_text_input, encode_images
pe1, po1, _ = prepare_text_input(src.flux_pipe, ["a cat"]); pe2, po2, _ = prepare_text_input(m.flux_pipe, ["a cat"])
print("text equal:", torch.equal(pe1, pe2), torch.equal(po1, po2))
torch.manual_seed(5); t1, _ = encode_images(src.flux_pipe, img); torch.manual_seed(5); t2, _ = encode_images(m.flux_pipe, img)
print("cond tokens equal:", torch.equal(t1, t2))
