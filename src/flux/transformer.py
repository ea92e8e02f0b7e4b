"""Drop-in for the reference's src/flux/transformer.py: `tranformer_forward` [sic] (transformer.py:47-252).

Same signature and return convention; the arithmetic is one native call sequence (lx_dit_prepare + lx_dit_step,
loongx_b200/csrc/engine.cu).  `generate()` does not call this per step — it prepares all timesteps once — but the
train-style call sites (model.py:705-723) and external callers keep working unchanged.
"""
from typing import Any, Dict, Optional

import torch


class Transformer2DModelOutput:
    def __init__(self, sample):
        self.sample = sample


def prepare_params(hidden_states, encoder_hidden_states=None, pooled_projections=None, timestep=None, img_ids=None,
                   txt_ids=None, guidance=None, joint_attention_kwargs=None, controlnet_block_samples=None,
                   controlnet_single_block_samples=None, return_dict=True, **kwargs):
    return (hidden_states, encoder_hidden_states, pooled_projections, timestep, img_ids, txt_ids, guidance,
            joint_attention_kwargs, controlnet_block_samples, controlnet_single_block_samples, return_dict)


def tranformer_forward(transformer, condition_latents, condition_ids, condition_type_ids,
                       model_config: Optional[Dict[str, Any]] = {}, c_t=0, **params):
    (hidden_states, encoder_hidden_states, pooled_projections, timestep, img_ids, txt_ids, guidance,
     joint_attention_kwargs, controlnet_block_samples, controlnet_single_block_samples, return_dict) = prepare_params(**params)
    # transformer.py:73-83, 246-248: the LoRA layers are scaled by joint_attention_kwargs["scale"] for this forward
    transformer.set_lora_scale(joint_attention_kwargs.get("scale", 1.0) if joint_attention_kwargs is not None else 1.0)
    if transformer.training and transformer.gradient_checkpointing:
        raise NotImplementedError("tranformer_forward is the inference forward; the differentiable training path is "
                                  "OminiModel.step (loongx_b200/train.py), which checkpoints / recomputes per block")
    model_config = model_config or {}
    use_condition = condition_latents is not None
    if txt_ids.ndim == 3:  # transformer.py:117-128 (deprecated batched ids)
        txt_ids = txt_ids[0]
    if img_ids.ndim == 3:
        img_ids = img_ids[0]
    B, n_img, _ = hidden_states.shape
    n_txt = encoder_hidden_states.shape[1]
    n_cond = condition_latents.shape[1] if use_condition else 0
    plan = transformer.plan(B, n_txt, n_img, n_cond, 1, model_config, transformer.c_factor())
    # condition_type_ids is ignored exactly like the reference (transformer.py:133 is commented out)
    plan.set_ids(txt_ids, img_ids, condition_ids if use_condition else None)
    ts = timestep.detach().float().reshape(-1).tolist()
    if len(ts) == 1 and B > 1:
        ts = ts * B
    gd = None
    if guidance is not None:
        gd = guidance.detach().float().reshape(-1).tolist()
        if len(gd) == 1 and B > 1:
            gd = gd * B
    plan.prepare(encoder_hidden_states, pooled_projections, condition_latents if use_condition else None, ts, gd,
                 c_t=float(c_t))
    if controlnet_block_samples is not None or controlnet_single_block_samples is not None:
        # transformer.py:172-181, 230-239: residuals on the image stream after every block -> block-by-block forward
        output = plan.step_with_residuals(0, hidden_states.to(torch.bfloat16).contiguous(), controlnet_block_samples,
                                          controlnet_single_block_samples).to(hidden_states.dtype)
    else:
        output = plan.step(0, hidden_states.to(torch.bfloat16).contiguous()).to(hidden_states.dtype)
    if not return_dict:
        return (output,)
    return Transformer2DModelOutput(sample=output)
