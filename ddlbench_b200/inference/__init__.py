from .beam_search import SequenceGenerator  # noqa: F401
