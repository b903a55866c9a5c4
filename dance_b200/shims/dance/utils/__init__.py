import os
import random

import numpy as np
import torch

from dance import logger


def get_device(device: str = "auto") -> str:
    if device == "auto":
        return "cuda" if torch.cuda.is_available() else "cpu"
    return device


def set_seed(rndseed, cuda: bool = True, extreme_mode: bool = False):
    os.environ["PYTHONHASHSEED"] = str(rndseed)
    random.seed(rndseed)
    np.random.seed(rndseed)
    torch.manual_seed(rndseed)
    if cuda and torch.cuda.is_available():
        torch.cuda.manual_seed_all(rndseed)
    logger.info(f"Setting global random seed to {rndseed}")
