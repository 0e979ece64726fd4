#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "radon_pair_kernel_variants or pr_256_mask_counts" 2>&1 | tail -8
