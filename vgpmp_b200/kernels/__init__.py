from .kernels import *  # noqa: F401,F403
