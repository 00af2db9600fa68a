set -x
python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "rotary or tree_decode or stress or sampled_oracle or transformer" 2>&1 | tail -25
python tools/gpu_dev_check.py --only perfdec --timeout 90 --log gpurun_out/dev_r2h.log 2>&1 | cut -c1-330
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/bench_configs.py --which decode --iters 20 2>&1 | grep -E "^\{|Error|error" | cut -c1-400
