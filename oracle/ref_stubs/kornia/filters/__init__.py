import torch


def blur_pool2d(input: torch.Tensor, kernel_size: int) -> torch.Tensor:
    """Typed no-op: generic_utils.pyrdown is TorchScript-compiled at import and needs a real signature."""
    return input
