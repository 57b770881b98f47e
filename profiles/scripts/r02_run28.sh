mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tcgen05.py -m gpu -q -k "attention or forward_golden" 2>&1 | tail -3
timeout 1200 python -m pytest tests/test_gpu_training.py -m gpu -q --tb=short -s 2>&1 | grep -v Warning | tail -70
