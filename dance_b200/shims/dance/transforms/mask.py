from dance_b200.transforms.mask import *  # noqa: F401,F403
