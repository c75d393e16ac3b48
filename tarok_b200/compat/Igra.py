from tarok_b200.igra import Igra, shuffle  # noqa: F401
