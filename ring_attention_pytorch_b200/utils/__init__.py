from ring_attention_pytorch_b200.utils.tensor_typing import Bool, Float, Int  # noqa: F401
from ring_attention_pytorch_b200.utils.timing import attention_flops, max_over_ranks, nvtx_range, time_cuda  # noqa: F401
