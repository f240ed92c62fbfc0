from .vgpmp import VGPMP, AdamConfig, initialize_Z  # noqa: F401
