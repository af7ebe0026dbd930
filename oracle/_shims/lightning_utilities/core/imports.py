class RequirementCache:
    """Truthy for torch/lightning pins, falsy for optional accelerators (flash-attn, bitsandbytes)
    so the reference takes its plain SDPA path."""

    def __init__(self, requirement: str, module: str = None) -> None:
        self.requirement = requirement

    def __bool__(self) -> bool:
        r = self.requirement.lower()
        return not (r.startswith("flash-attn") or r.startswith("bitsandbytes") or r.startswith("torch_xla"))

    def __str__(self) -> str:
        return f"Requirement {self.requirement!r} stubbed"
