#!/bin/bash
timeout 900 python -m pytest tests/test_cli.py tests/test_gpu_binvox.py tests/test_reference_main.py -x -q 2>&1 | tail -4
