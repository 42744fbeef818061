from .encoder import ConformerEncoder  # noqa: F401
