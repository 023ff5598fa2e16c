#!/bin/bash
python -m pytest tests/test_gpu_instancing.py tests/test_gpu_traversal.py tests/test_gpu_render.py -m gpu -q 2>&1 | tail -40
