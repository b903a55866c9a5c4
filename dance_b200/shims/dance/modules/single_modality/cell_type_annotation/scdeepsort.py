from dance_b200.modules.scdeepsort import ScDeepSort  # noqa: F401
