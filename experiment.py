#!/usr/bin/env python
"""python experiment.py --config dafnet_config_chaos --split 0 --l_mix 1   (reference: experiment.py:127-129)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from multimodal_segmentation_b200.experiment import Experiment  # noqa: E402

if __name__ == "__main__":
    Experiment().run()
