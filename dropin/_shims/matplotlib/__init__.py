"""Import-only stand-in for matplotlib (dataset/gta5_dataset.py:5 imports pyplot and never uses it)."""
