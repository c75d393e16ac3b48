from tarok_b200.igra import Tarok  # noqa: F401
