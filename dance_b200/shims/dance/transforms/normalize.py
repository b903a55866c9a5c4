from dance_b200.transforms.normalize import *  # noqa: F401,F403
