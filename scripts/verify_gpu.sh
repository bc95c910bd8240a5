set -x
python -m pytest tests/test_models_gpu.py -x -q -s -k "dice_within" 2>&1 | tail -8
