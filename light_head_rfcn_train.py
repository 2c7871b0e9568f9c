"""Repo-root launcher with the reference's script name: ``python light_head_rfcn_train.py [--flags]``."""
import xdet_b200  # noqa: F401
from xdet_b200.light_head_rfcn_train import main

if __name__ == "__main__":
    main()
