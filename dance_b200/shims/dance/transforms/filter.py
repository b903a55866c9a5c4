from dance_b200.transforms.filter import *  # noqa: F401,F403
