"""Committed SASS evidence: per kernel, the counts of the Blackwell-native mnemonics and short excerpts around the
interesting instruction groups (tcgen05.mma batches, TMA loads, bulk copies of the NVLink fetch, TMA reductions,
multimem loads, TMEM loads / stores, MUFU chunks).  Runs without a GPU:

    python tools/sass_excerpts.py            # reads ring_attention_pytorch_b200/_build/*.o, writes profiles/sass/
"""
from __future__ import annotations

import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "ring_attention_pytorch_b200", "_build")
OUT = os.path.join(ROOT, "profiles", "sass")

# PTX -> SASS names that prove the native path (B200_PROFILING.md): tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM,
# cp.async.bulk.tensor -> UTMALDG, cp.async.bulk -> UBLKCP, cp.reduce.async.bulk.tensor -> UTMAREDG,
# multimem.ld_reduce -> LDGMC, tcgen05.commit -> UTCBAR, mbarrier try_wait -> SYNCS.PHASECHK
INTEREST = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMAREDG", "UBLKCP", "UBLKRED", "LDGMC", "LDTM", "STTM", "UTCBAR", "MUFU.EX2",
            "FFMA2", "HMMA"]
EXCERPT = ["UTCHMMA", "UTCQMMA", "UTMAREDG", "UBLKCP", "LDGMC", "MUFU.EX2"]
CTX = 14
# excerpts are written for one representative instantiation per kernel (bf16, head dim 128); SUMMARY.txt lists all
WRITE = ("attn_fwd_kernel<128,true,1>", "attn_bwd_fused_kernel<true,true,false>", "attn_bwd_fused_kernel<true,false,false>",
         "attn_bwd_dq_kernel<128,true>", "attn_bwd_dkdv_kernel<128,true>", "tree_decode_tc_kernel<false,0>",
         "tree_decode_tc_kernel<true,0>", "tree_decode_kernel<128,0>")


def demangle(name: str) -> str:
    try:
        out = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        return out or name
    except Exception:  # noqa: BLE001
        return name


def short(name: str) -> str:
    m = re.search(r"(\w+_kernel)<([^>]*)>", name)
    if m:
        return f"{m.group(1)}<{m.group(2).replace(' ', '')}>"
    m = re.search(r"(\w+_kernel)", name)
    return m.group(1) if m else name[:60]


def main():
    os.makedirs(OUT, exist_ok=True)
    summary = []
    for obj in sorted(os.listdir(BUILD)):
        if not obj.endswith(".cu.o"):
            continue
        sass = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, obj)], capture_output=True, text=True).stdout
        funcs = re.split(r"\n\s*Function : ", sass)[1:]
        for f in funcs:
            head, _, body = f.partition("\n")
            name = short(demangle(head.strip()))
            lines = [ln for ln in body.split("\n") if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln)]
            ops = Counter()
            for ln in lines:
                m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
                if m:
                    op = m.group(1)
                    for key in INTEREST:
                        if op.startswith(key):
                            ops[key] += 1
            if not any(ops[k] for k in ("UTCHMMA", "UTCQMMA", "UTMALDG", "UBLKCP", "LDGMC", "UTMAREDG")):
                continue
            counts = ", ".join(f"{k} {ops[k]}" for k in INTEREST if ops[k])
            summary.append(f"{obj[:-5]:28s} {name:60s} instr {len(lines):6d} | {counts}")
            if name not in WRITE:
                continue
            safe = re.sub(r"[^A-Za-z0-9_]+", "_", name)[:80]
            with open(os.path.join(OUT, f"{obj[:-5]}__{safe}.sass.txt"), "w") as out:
                out.write(f"// {name}\n// {len(lines)} SASS instructions; {counts}\n")
                for key in EXCERPT:
                    idx = [i for i, ln in enumerate(lines) if re.search(r"\*/\s+(?:@!?U?P\d+\s+)?" + re.escape(key), ln)]
                    if not idx:
                        continue
                    # the densest window: most occurrences of `key` within CTX * 2 lines
                    best = max(idx, key=lambda i: sum(1 for j in idx if i <= j < i + 2 * CTX))
                    lo, hi = max(0, best - 4), min(len(lines), best + 2 * CTX)
                    out.write(f"\n// ---- {key}: {len(idx)} occurrences; densest window (instructions {lo}..{hi}) ----\n")
                    out.write("\n".join(ln.rstrip() for ln in lines[lo:hi]) + "\n")
    with open(os.path.join(OUT, "SUMMARY.txt"), "w") as f:
        f.write("object                       kernel                                                       size   | native mnemonics\n")
        f.write("\n".join(summary) + "\n")
    print("\n".join(summary))


if __name__ == "__main__":
    sys.exit(main())
