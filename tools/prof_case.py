"""One forward + backward of the fused kernels on a mid-sized problem, for ncu captures:

    ncu --set full --clock-control none --import-source on -k regex:attn_ -c 3 -o gpurun_out/prof \
        python tools/prof_case.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ring_attention_pytorch_b200.ops import _ext  # noqa: E402
from ring_attention_pytorch_b200.ops.fused import (alloc_kv_buffer, alloc_qdo_buffer, alloc_stat_buffer,  # noqa: E402
                                                   fused_attn_bwd, fused_attn_fwd)
from ring_attention_pytorch_b200.parallel.layout import make_position_map  # noqa: E402

n = int(os.environ.get("PROF_N", 8192))
h = int(os.environ.get("PROF_H", 8))
d = int(os.environ.get("PROF_D", 128))
causal = os.environ.get("PROF_CAUSAL", "1") == "1"
ops = _ext.ops()
dt = torch.bfloat16
q, k, v, do = (torch.randn(1, n, h, d, device="cuda", dtype=dt) for _ in range(4))
pm = make_position_map("plain", 1, n)
kv = alloc_kv_buffer(1, 1, h, n, d, dt, "cuda")
qdo = alloc_qdo_buffer(1, 1, h, n, d, dt, "cuda")
stat = alloc_stat_buffer(1, 1, h, n, "cuda")
ready = torch.zeros(1, dtype=torch.int32, device="cuda")
ops.pack_kv(k, v, kv[0])
o, lse = fused_attn_fwd(q, kv, [0], ready, None, kv_heads=h, rank=0, pm=pm, causal=causal, window=None, scale=d ** -0.5)
ops.bwd_prep(q, o, do, lse, qdo, stat, 0)
fused_attn_bwd(qdo, kv, stat, None, batch=1, heads=h, kv_heads=h, rank=0, pm=pm, causal=causal, window=None,
               scale=d ** -0.5)
torch.cuda.synchronize()
print("done")
