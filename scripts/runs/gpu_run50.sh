set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 4
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-300
