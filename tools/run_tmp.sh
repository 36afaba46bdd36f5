# scratch script for gpurun calls during development (overwritten freely)
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
