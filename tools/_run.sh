set -x
python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "tree_decode" 2>&1 | tail -8
python tools/gpu_dev_check.py --only perfdec --timeout 90 --log gpurun_out/dev_r2g.log 2>&1 | cut -c1-300
cuobjdump -sass ring_attention_pytorch_b200/_C.so | grep -c MULTIMEM
