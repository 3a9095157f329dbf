#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -x -q -k lstm 2>&1 | tail -25
timeout 600 python tools/gpu_profile_train.py bf16 lstm 2>&1 | tail -40
