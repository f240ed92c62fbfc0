from .vgpmp import VGPMP, AdamConfig, initialize_Z  # noqa: F401
from .streamed import StreamedVGPMP  # noqa: F401
