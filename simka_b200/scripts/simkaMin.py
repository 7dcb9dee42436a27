#!/usr/bin/env python
"""simkaMin.py -- same command line as the reference's simkaMin/simkaMin.py, driving this build's simkaMinCore (GPU)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from simka_b200.simkamin_pipeline import main_simkamin  # noqa: E402

if __name__ == "__main__":
    sys.exit(main_simkamin())
