"""Drop-in for the reference's synthetic_static_obs/validation.py (same CLI, reads ./data/*.npz, writes ./stats/*.npz) with the 1000-rollout
Monte-Carlo evaluation of every planned trajectory on the GPU (mpcmmd_b200/validation.py, csrc/k_validate.cuh)."""
import os
import sys

sys.path.insert(1, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mpcmmd_b200.validation import main  # noqa: E402

if __name__ == "__main__":
    main(variant="static")
