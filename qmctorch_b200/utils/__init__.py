"""Precision switches, same names as qmctorch/utils/torch_utils.py:8-21."""
import torch


def set_torch_double_precision():
    """qmctorch/utils/torch_utils.py:8-13 - the CUDA path computes in FP64 only."""
    torch.set_default_dtype(torch.float64)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def set_torch_single_precision():
    raise NotImplementedError(
        "qmctorch_b200 evaluates the wave function in FP64 only (BASELINE north_star); "
        "single precision (qmctorch/utils/torch_utils.py:16-21) is not provided")
