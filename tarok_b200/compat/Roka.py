from tarok_b200.karte import Roka  # noqa: F401
