"""Drop-in for the reference's synthetic_static_obs/main_mpc.py (same CLI, same ./data/*.npz schema) on the batched B200 solver."""
import os
import sys

sys.path.insert(1, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mpcmmd_b200.driver import main  # noqa: E402

if __name__ == "__main__":
    main(variant="static")
