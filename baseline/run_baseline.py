"""Run the reference-equivalent stock-PyTorch arm (see baseline/README.md)."""
import os
import runpy
import sys

if __name__ == "__main__":
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.environ["ZTB_FORCE_REFERENCE"] = "1"
    sys.argv = [os.path.join(root, "bench.py"), "--impl", "baseline", *sys.argv[1:]]
    runpy.run_path(os.path.join(root, "bench.py"), run_name="__main__")
