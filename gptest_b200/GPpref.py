"""Drop-in replacement for the reference's ``GPpref.py`` (under construction in this commit)."""
