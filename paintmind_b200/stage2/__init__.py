from .transformer import CondTransformer  # noqa: F401
