from .likelihood import *  # noqa: F401,F403
