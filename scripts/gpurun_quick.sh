#!/bin/bash
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_mesh.py tests/test_gpu_multi.py tests/test_gpu_readback.py tests/test_gpu_binvox.py -x -q 2>&1 | tail -2
