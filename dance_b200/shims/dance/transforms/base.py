from dance_b200.transforms.base import *  # noqa: F401,F403
