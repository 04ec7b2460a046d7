"""tools/train_steps.py with set_train_precision('tf32') (profiling entry point)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import snerf_b200  # noqa: E402
import train_steps  # noqa: E402

snerf_b200.set_train_precision("tf32")
train_steps.main()
