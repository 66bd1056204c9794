set -x
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 4
