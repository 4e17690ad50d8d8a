#!/usr/bin/env python
"""Same command line as the reference's UnMicstCyto2.py:679-827, running on the B200 engine (unmicst_b200.cli)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from unmicst_b200.cli import run_tool, run_wrapper  # noqa: E402,F401

if __name__ == "__main__":
    sys.exit(run_tool('UnMicstCyto2'))
