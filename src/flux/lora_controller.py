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
    """lora_controller.py:45-75.  A LoRA scale other than 1 would need re-merging W + s (alpha/r) B A."""

    def __init__(self, lora_modules: List[Any], scale: float) -> None:
        if scale != 1:
            raise NotImplementedError("set_lora_scale(scale != 1): LoRA is merged into the condition-row weight panel at load")
        self.lora_modules = list(lora_modules)
        self.scale = scale

    def __enter__(self) -> None:
        return None

    def __exit__(self, exc_type, exc_val, exc_tb) -> None:
        return None
