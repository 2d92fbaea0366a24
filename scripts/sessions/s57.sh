#!/bin/bash
D=gpurun_out/s57; mkdir -p $D
timeout 600 python -m pytest tests/test_gpu_reader2.py -m gpu -x -q -k registered 2>&1 | tail -2
python scripts/probe_register_sequence.py 2>&1 | tail -9 | tee $D/register_sequence.txt
