#!/usr/bin/env python
"""simkaMin_update.py -- same command line as the reference's simkaMin/simkaMin_update.py, driving this build's simkaMinCore (GPU)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from simka_b200.simkamin_pipeline import main_update  # noqa: E402

if __name__ == "__main__":
    sys.exit(main_update())
