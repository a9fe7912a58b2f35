"""Sampling entry point with the reference's name:
``python app.py --model-size 417m --model-path out.pth [--share]`` (reference ``app.py:14-20``)."""
from zero_transformer_b200.generate import main

if __name__ == "__main__":
    raise SystemExit(main())
