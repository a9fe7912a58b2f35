"""Entry point with the reference's name and flags:
``python main_zero.py [--cfg conf/config.yaml] [--model-cfg conf/model_config.yaml] [--resume]``
(reference ``main_zero.py:41-55``).  Multi-GPU: ``torchrun --nproc-per-node N main_zero.py ...``."""
from zero_transformer_b200.train import main

if __name__ == "__main__":
    raise SystemExit(main())
