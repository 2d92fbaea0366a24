#!/bin/bash
timeout 400 python -m pytest tests/test_inflate.py -m gpu -x -q 2>&1 | tail -15
