from tarok_b200.igra import Berac  # noqa: F401
