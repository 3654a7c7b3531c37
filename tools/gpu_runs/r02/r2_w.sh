#!/bin/bash
# round 2, call W: the new bit-identity test of the fused RoPE epilogue
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 700 -p no:cacheprovider -k "rope_epilogue" 2>&1 | tail -15
