from tarok_b200.karte import Barva, Karta  # noqa: F401
