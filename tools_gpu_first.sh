#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "abc or text_3k or run260 or tiny" --tb=short 2>&1 | tail -80 > gpurun_out/sanitizer.log
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider --tb=short 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
echo ==== sanitizer; tail -40 gpurun_out/sanitizer.log
echo ==== pytest; cat gpurun_out/pytest_gpu.log
