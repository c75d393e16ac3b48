from tarok_b200.igra import Navadna_igra  # noqa: F401
