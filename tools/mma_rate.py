#!/usr/bin/env python
"""tcgen05.mma issue-rate calibration (``torch.ops.rab.umma_rate``): cycles per M=128 x N x K=16 instruction for the
operand flavours the attention kernels use, on 1 SM and on all 148 SMs at once.  Writes gpurun_out/mma_rate.json.

Ideal (tensor-bound) cost is N/2 cycles per instruction (8192 dense bf16 FLOP/clk/SM)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ring_attention_pytorch_b200.ops import _ext  # noqa: E402

MODES = {0: "SS  B K-major", 1: "SS  B MN-major", 2: "TS  B MN-major", 3: "TS  B K-major"}


def main():
    ops = _ext.ops()
    reps = 4096
    rows = []
    for ctas in (148,):
        for mode in MODES:
            for n in (64, 128, 256):
                for alt in (0, 1):
                    ops.umma_rate(mode, n, 256, alt, ctas)  # warm-up
                    torch.cuda.synchronize()
                    out = ops.umma_rate(mode, n, reps, alt, ctas).cpu()
                    total = out[:, 0].float().mean().item() / reps
                    issue = out[:, 1].float().mean().item() / reps
                    rows.append(dict(ctas=ctas, mode=MODES[mode], n=n, alt=alt, cyc_per_mma=round(total, 1),
                                     issue_cyc_per_mma=round(issue, 1), ideal=n / 2,
                                     efficiency=round(n / 2 / total, 3)))
                    print(rows[-1], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "mma_rate.json"), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
