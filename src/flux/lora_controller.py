"""Drop-in for the reference's src/flux/lora_controller.py (enable_lora / set_lora_scale, lora_controller.py:5-75).

In the reference these context managers zero peft's `scaling` on the listed modules so that LoRA acts on the
condition branch only.  In this build that masking is structural: every LoRA-targeted Linear keeps a base panel W and a
merged panel W + (alpha/r) B A, and the GEMM's row groups decide which rows read which panel
(loongx_b200/csrc/engine.cu).  The context managers therefore only record the request; `activated=True` (the reference's
`latent_lora`) is honoured through `model_config["latent_lora"]`, which selects the merged panel for the image rows.
"""
from typing import Any, List, Optional, Type


class enable_lora:
    def __init__(self, lora_modules: List[Any], activated: bool) -> None:
        self.activated = activated
        self.lora_modules = list(lora_modules)

    def __enter__(self) -> None:
        return None

    def __exit__(self, exc_type: Optional[Type[BaseException]], exc_val: Optional[BaseException], exc_tb: Optional[Any]) -> None:
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
        seen, self._transformers = set(), []
        for m in self.lora_modules:
            tr = m if hasattr(m, "set_lora_scale") and hasattr(m, "weights") else getattr(m, "transformer", None)
            if tr is None and hasattr(m, "block"):
                tr = getattr(m.block, "transformer", None)
            if tr is not None and id(tr) not in seen:
                seen.add(id(tr))
                self._transformers.append(tr)
        self.scales = [float(getattr(tr.weights, "lora_scale", 1.0)) for tr in self._transformers]

    def __enter__(self) -> None:
        for tr, prev in zip(self._transformers, self.scales):
            tr.set_lora_scale(prev * self.scale)
        return None

    def __exit__(self, exc_type: Optional[Type[BaseException]], exc_val: Optional[BaseException], exc_tb: Optional[Any]) -> None:
        for tr, prev in zip(self._transformers, self.scales):
            tr.set_lora_scale(prev)
        return None
