#!/bin/bash
# A/B of the given library tags (see gpu_ab2.sh), then the GPU test suite on the in-tree build
bash scripts/gpu_ab2.sh "$@"
( time timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 ) 2>&1 | cut -c1-200
