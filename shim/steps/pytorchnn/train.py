#!/usr/bin/env python
"""Drop-in for ``python steps/pytorchnn/train.py ...`` (stage 1 of run_nnlm_{ami,lrs2}_{lstm,tm}.sh): the reference
trainer's command line, corpus files, schedule and checkpoint format, with the step running on bayeslms_b200."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "..")))

from bayeslms_b200.train import main  # noqa: E402

if __name__ == "__main__":
    raise SystemExit(main())
