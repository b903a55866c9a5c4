from dance_b200.matrix import *  # noqa: F401,F403
