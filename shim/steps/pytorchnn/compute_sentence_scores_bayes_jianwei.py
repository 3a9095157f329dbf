#!/usr/bin/env python
"""Drop-in for stage 6 of lmrescore_nbest_pytorchnn_cuda.sh (lines 197-219): same command line, same
``words_text`` -> ``lmwt.nn`` files as the reference scorer of this name, scored by bayeslms_b200 on a B200."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "..")))

from bayeslms_b200.scorer import main  # noqa: E402

if __name__ == "__main__":
    raise SystemExit(main())
