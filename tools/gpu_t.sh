#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "$1" 2>&1 | tail -25
