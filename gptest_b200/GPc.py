"""Drop-in replacement for the reference's ``GPc.py`` (under construction in this commit)."""
