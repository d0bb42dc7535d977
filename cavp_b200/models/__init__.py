from .cavp_model import CAVP, SoundBank  # noqa: F401
