#!/bin/bash
# First GPU call of the next round: exercises everything that was written after the last GPU minute of round 1.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/next_round_first_call.sh'
# Every step has its own timeout and writes into gpurun_out/; nothing here changes defaults.
mkdir -p gpurun_out
log=gpurun_out/first_call.log
: > "$log"
run() { echo "==== $*" | tee -a "$log"; ( "$@" ) 2>&1 | tail -40 | tee -a "$log"; }

# 1. operand flavours for the one-kernel backward (probe mode 3)
run timeout 60 python tools/gpu_dev_check.py --only xprobe_mn_a,xprobe_mn_a_k64 --timeout 25
# 2. split-half forward: numerics first, then perf against the default
RAB_FWD_SPLIT=1 run timeout 200 python tools/gpu_dev_check.py --timeout 40 \
    --only fwd_d128_causal_n1000,fwd_many_items,fwd_d64_causal_n777,fwd_kmask,ring4_striped_causal,ring3_kmask,perf_causal_16k,perf_causal_64k_h8
run timeout 60 python tools/gpu_dev_check.py --timeout 40 --only perf_causal_16k,perf_causal_64k_h8
# 3. FMA-pipe exponentials in the backward
RAB_BWD_EXP_POLY=1 run timeout 120 python tools/gpu_dev_check.py --timeout 40 \
    --only bwd_d128_causal_n1000,bwd_many_items,bwd_d64_causal_n777,perfbwd_causal_16k,perfbwd_causal_64k_h8
# 3b. experimental one-kernel backward (needs step 1 green: it relies on the probe-mode-3 operand flavours)
run timeout 120 python tools/gpu_dev_check.py --timeout 50 --only xfused_bwd_causal,xfused_bwd_full_gqa
# 4. bench with the prefetching e2e loop, then the reference arm with its probe
run timeout 240 python bench.py --steps 3 --warmup 3
run timeout 600 python bench.py --impl reference --steps 2 --warmup 3
