set -x
python -m pytest tests/test_models_gpu.py -x -q -k "predict_mask" 2>&1 | tail -8
