from tarok_b200.karte import Tip_igre  # noqa: F401
