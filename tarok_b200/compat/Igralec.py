from tarok_b200.igralec import Bot_igralec, Igralec  # noqa: F401
