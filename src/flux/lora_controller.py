"""Drop-in for the reference's src/flux/lora_controller.py (enable_lora / set_lora_scale, lora_controller.py:5-75).

In the reference these context managers zero peft's `scaling` on the listed modules so that LoRA acts on the
condition branch only.  In this build that masking is structural: every LoRA-targeted Linear keeps a base panel W and a
merged panel W + (alpha/r) B A, and the GEMM's row groups decide which rows read which panel
(loongx_b200/csrc/engine.cu); the reference's `latent_lora` (`activated=True` inside its block code) is honoured through
`model_config["latent_lora"]`, which selects the merged panel for the image rows.  Called by a user around native calls,
both context managers act on the merged panels of the whole weight set (see the classes).
"""
from typing import Any, List, Optional, Type


def _native_transformers(lora_modules: List[Any]) -> List[Any]:
    """the native transformers the listed handles belong to (the transformer itself, one of its block handles, or a
    block's `.attn`); anything else is skipped, like the reference skips modules that are not peft `BaseTunerLayer`s"""
    seen, out = set(), []
    for m in lora_modules:
        tr = m if hasattr(m, "set_lora_outer") and hasattr(m, "weights") else getattr(m, "transformer", None)
        if tr is None and hasattr(m, "block"):
            tr = getattr(m.block, "transformer", None)
        if tr is not None and id(tr) not in seen:
            seen.add(id(tr))
            out.append(tr)
    return out


class enable_lora:
    """lora_controller.py:5-43: with `activated=False` the LoRA scaling of the listed modules is zero inside the `with`
    block and restored afterwards; `activated=True` leaves it alone.  The reference uses it INSIDE its block code to keep
    LoRA off the text / image rows - here that is structural (see the module docstring) and needs no call.  Used by a
    caller AROUND native calls it means what it means in the reference for everything computed inside: LoRA off, i.e.
    the merged panels of the transformer(s) the handles belong to are rebuilt at scale 0 (W + 0 B A = W exactly) and put
    back at the previous scale on exit."""

    def __init__(self, lora_modules: List[Any], activated: bool) -> None:
        self.activated = activated
        self.lora_modules = list(lora_modules)
        self._transformers = [] if activated else _native_transformers(self.lora_modules)
        self.scales: List[float] = []

    def __enter__(self) -> None:
        self.scales = [float(getattr(tr.weights, "lora_outer", 1.0)) for tr in self._transformers]
        for tr in self._transformers:
            tr.set_lora_outer(0.0)
        return None

    def __exit__(self, exc_type: Optional[Type[BaseException]], exc_val: Optional[BaseException], exc_tb: Optional[Any]) -> None:
        for tr, prev in zip(self._transformers, self.scales):
            tr.set_lora_outer(prev)
        return None


class set_lora_scale:
    """lora_controller.py:45-75: multiply the LoRA scaling of the listed modules by `scale` inside the `with` block and
    restore it afterwards.  Native granularity: the handles generate() / block.py see (`transformer.transformer_blocks[i]`,
    `.attn`, or the transformer itself) stand for whole blocks of ONE weight set whose merged panels
    W + s (alpha/r) B A are rebuilt by the native merge kernel, so the scale applies to every LoRA target of the
    transformer(s) the listed modules belong to; objects that are not native handles are skipped, like the reference
    skips modules that are not peft `BaseTunerLayer`s.  (The reference never calls this class; it is API surface.)"""

    def __init__(self, lora_modules: List[Any], scale: float) -> None:
        self.lora_modules = list(lora_modules)
        self.scale = float(scale)
        self._transformers = _native_transformers(self.lora_modules)
        self.scales: List[float] = []

    def __enter__(self) -> None:
        # the multiplier lives next to (not in place of) the per-forward joint_attention_kwargs scale, which
        # tranformer_forward sets on every call: like peft's scale_lora_layers, the two compose
        self.scales = [float(getattr(tr.weights, "lora_outer", 1.0)) for tr in self._transformers]
        for tr, prev in zip(self._transformers, self.scales):
            tr.set_lora_outer(prev * self.scale)
        return None

    def __exit__(self, exc_type: Optional[Type[BaseException]], exc_val: Optional[BaseException], exc_tb: Optional[Any]) -> None:
        for tr, prev in zip(self._transformers, self.scales):
            tr.set_lora_outer(prev)
        return None
