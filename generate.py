#!/usr/bin/env python
"""`python generate.py ...` - the reference's CLI (reference generate.py:69-247) on gtav_b200's CUDA path.
See ai-generated-gtav_b200/generate.py for the flags and the documented differences."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from gtav_b200.generate import main  # noqa: E402

if __name__ == "__main__":
    raise SystemExit(main())
