from dance_b200.transforms.interface import *  # noqa: F401,F403
