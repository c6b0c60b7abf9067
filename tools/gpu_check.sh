#!/bin/bash
# Runs on the GPU box (gpurun): smoke, the -m gpu parity tests and a short bench.
# Everything interesting lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== smoke" 
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -15 | tee gpurun_out/smoke.log
echo "== pytest -m gpu"
timeout ${PYTEST_TIMEOUT:-1500} python -m pytest tests -m gpu -q --tb=short -x --timeout=300 ${PYTEST_ARGS} 2>&1 | tail -${PYTEST_TAIL:-60} | tee gpurun_out/pytest_gpu.log
