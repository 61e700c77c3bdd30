"""Drop-in mirror of the reference's `src.defenses` package for the purification path."""
