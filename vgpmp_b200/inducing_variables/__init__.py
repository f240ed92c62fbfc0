from .inducing_variables import *  # noqa: F401,F403
