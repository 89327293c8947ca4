"""Kept for import compatibility (reference ``gto/_init_paths.py`` only edits ``sys.path``)."""
